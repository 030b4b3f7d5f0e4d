"""Persistent chain kernel (TT_CHAIN=1) against the per-kernel chain (TT_CHAIN=0) on the same inputs: loss, h, stats,
all 8 gradients (+ table gradients), and the time of the back phase of a step.  GPU box only."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import two_towers_oracle as O  # noqa: E402
from two_towers_overlords_b200 import TwoTowersModel  # noqa: E402
from two_towers_overlords_b200.training import FusedTrainer  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def run(B, P, precision, train_table, chain, shape="Z", iters=20):
    os.environ["TT_CHAIN"] = str(chain)
    torch.manual_seed(0)
    m = TwoTowersModel(projection_dim=P, precision=precision, train_table=train_table).cuda()
    batch = O.synth_triplet_batch(B, 32, 256, shape, seed=5)
    Lq, Ld = batch.q_ids.shape[1], batch.p_ids.shape[1]
    tr = FusedTrainer(m, 0.3, 1e-3, B, Lq, Ld, precision=precision, use_graph=False, ids_dtype=torch.int32,
                      mask_dtype=torch.uint8)
    for dst, src in zip(tr.tok, batch.astuple()):
        dst.copy_(src.to(dst.dtype))
    tr._fwd_bwd()
    torch.cuda.synchronize()
    v = tr.step_obj._views()
    out = {"loss": tr.loss_view.clone(), "h": tr.step_obj.hidden().clone(), "stats": v["stats"].clone(),
           "grads": [g.clone() for g in tr.g_views]}
    if train_table:
        out["tables"] = [g.clone() for g in tr.table_grads]
        out["dxhat"] = v["dxhat"].clone()
    # time the back phase alone (gather already done)
    tr._fwd_bwd(0, 1, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        tr._fwd_bwd(0, 2, 0)
    e0.record()
    for _ in range(iters):
        tr._fwd_bwd(0, 2, 0)
    e1.record()
    torch.cuda.synchronize()
    out["back_us"] = e0.elapsed_time(e1) / iters * 1e3
    # twice the same bits
    tr._fwd_bwd()
    torch.cuda.synchronize()
    out["repeat"] = all(torch.equal(a, b) for a, b in zip(out["grads"], tr.g_views))
    return out


cases = [(2048, 512, "bf16x3", False), (300, 64, "bf16x3", False), (256, 384, "bf16x3", True), (2048, 512, "bf16", False),
         (1000, 128, "bf16x3", False)]
if len(sys.argv) > 1:
    cases = [(int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], len(sys.argv) > 4 and sys.argv[4] == "table")]
for B, P, prec, tt_ in cases:
    t0 = time.time()
    a = run(B, P, prec, tt_, 0)
    b = run(B, P, prec, tt_, 1)
    print(f"--- B={B} P={P} {prec} table={tt_}: loss {float(a['loss']):.8f} / {float(b['loss']):.8f}  "
          f"back phase {a['back_us']:.1f} us (kernels) vs {b['back_us']:.1f} us (chain)  repeatable {b['repeat']}  "
          f"[{time.time() - t0:.1f} s]")
    print(f"  h {rel(b['h'], a['h']):.2e}  stats {rel(b['stats'], a['stats']):.2e}  gate flips "
          f"{int(((a['h'] > 0) != (b['h'] > 0)).sum())}")
    for i, (ga, gb) in enumerate(zip(a["grads"], b["grads"])):
        print(f"  grad[{i}] {rel(gb, ga):.2e}  |g| {float(ga.norm()):.3e}")
    if tt_:
        print(f"  dxhat {rel(b['dxhat'], a['dxhat']):.2e}  tables {rel(b['tables'][0], a['tables'][0]):.2e} "
              f"{rel(b['tables'][1], a['tables'][1]):.2e}")
