"""Where does the fused step's error against the CPU oracle come from?  Runs ONE step at configs[1] (B=2048, P=512)
per precision and prints the relative error of every intermediate (pooled rows, h, y, dY, dz1) per row group
(q | p | n) and of all 8 gradients.  GPU box only:  python scripts/diag_step_parity.py [shape] [B] [P]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import two_towers_oracle as O  # noqa: E402
from two_towers_overlords_b200 import TwoTowersModel  # noqa: E402
from two_towers_overlords_b200.training import FusedTrainer  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "U"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
P = int(sys.argv[3]) if len(sys.argv) > 3 else 512
H, margin = 384, 0.3


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def carve(ws, sizes):
    """Replicates carve_step_ws (tt_api.cu): consecutive 256-byte aligned fp32 arrays."""
    out, off = [], 0
    for n in sizes:
        off = (off + 255) // 256 * 256
        out.append(ws[off: off + n * 4].view(torch.float32))
        off += n * 4
    return out


batch = O.synth_triplet_batch(B, 32, 256, shape, seed=5)
for precision in sys.argv[4:] or ["fp32", "bf16x3"]:
    torch.manual_seed(0)
    m = TwoTowersModel(projection_dim=P, precision=precision).cuda()
    Lq, Ld = batch.q_ids.shape[1], batch.p_ids.shape[1]
    tr = FusedTrainer(m, margin, 1e-3, B, Lq, Ld, precision=precision, use_graph=False, ids_dtype=torch.int32,
                      mask_dtype=torch.uint8)
    for dst, src in zip(tr.tok, batch.astuple()):
        dst.copy_(src.to(dst.dtype))
    tr._fwd_bwd()
    torch.cuda.synchronize()
    R = 3 * B
    xhat, cnt, nrm, h, y, stats, dy, dz1 = carve(tr.step_obj.ws, [R * H, R, R, R * P, R * P, B * 8, R * P, R * P])
    ref = O.OracleTwoTowers(P)
    with torch.no_grad():
        for t_new, t_ref in ((m.query_tower, ref.query_tower), (m.document_tower, ref.document_tower)):
            t_ref.table.copy_(t_new.pretrained_model.table.cpu())
            for i in (0, 2):
                t_ref.projection[i].weight.copy_(t_new.projection[i].weight.cpu())
                t_ref.projection[i].bias.copy_(t_new.projection[i].bias.cpu())
    ref = ref.double()
    xs, hs, zs, ys = [], [], [], []
    for tower, ids, mask in ((ref.query_tower, batch.q_ids, batch.q_mask), (ref.document_tower, batch.p_ids, batch.p_mask),
                             (ref.document_tower, batch.n_ids, batch.n_mask)):
        x = torch.nn.functional.normalize(O.mean_pooling(torch.nn.functional.embedding(ids, tower.table), mask), dim=1)
        z = torch.nn.functional.linear(x, tower.projection[0].weight, tower.projection[0].bias)
        z.retain_grad()
        hh = torch.relu(z)
        yy = torch.nn.functional.linear(hh, tower.projection[2].weight, tower.projection[2].bias)
        yy.retain_grad()
        xs.append(x); zs.append(z); hs.append(hh); ys.append(yy)
    loss = O.triplet_loss(ys[0], ys[1], ys[2], margin)
    loss.backward()
    print(f"--- {precision} shape {shape} B={B} P={P}: loss {float(tr.loss_view.item()):.8f} vs {loss.item():.8f} "
          f"rel {abs(float(tr.loss_view.item()) - loss.item()) / loss.item():.2e}")
    for name, got, want in (("xhat", xhat.view(R, H), xs), ("h", h.view(R, P), hs), ("y", y.view(R, P), ys),
                            ("dY", dy.view(R, P), [t.grad for t in ys]), ("dz1", dz1.view(R, P), [t.grad for t in zs])):
        errs = [rel(got[g * B: (g + 1) * B], want[g]) for g in range(3)]
        print(f"  {name:5s} q {errs[0]:.2e}  p {errs[1]:.2e}  n {errs[2]:.2e}")
    # sign agreement of the ReLU gate
    for g, nm in enumerate("qpn"):
        flips = int(((h.view(R, P)[g * B: (g + 1) * B].cpu() > 0) != (zs[g] > 0)).sum())
        print(f"  relu sign flips {nm}: {flips} of {B * P}")
    names = [f"{t}.projection.{i}.{k}" for t in ("query_tower", "document_tower") for i in (0, 2) for k in ("weight", "bias")]
    pr = dict(ref.named_parameters())
    for nm, got in zip(names, tr.g_views):
        print(f"  grad {nm:40s} {rel(got, pr[nm].grad):.2e}   |g| {float(pr[nm].grad.norm()):.3e}")
    del tr, m
