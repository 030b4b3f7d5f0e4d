"""configs[2] (trainable tables): where does the table-gradient error against the oracle come from?  One step at
B = 4096, P = 384, shape Z; prints the relative error of dxhat and of both table gradients (all rows / rows named by
more than 256 tokens / the rest) for the device AND for the fp32 CPU oracle, both against a float64 oracle
differentiated through the device's ReLU gate.  GPU box only:  python scripts/diag_table_grad.py [B] [P] [precisions..]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import two_towers_oracle as O  # noqa: E402
from two_towers_overlords_b200 import TwoTowersModel  # noqa: E402
from two_towers_overlords_b200.training import FusedTrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
P = int(sys.argv[2]) if len(sys.argv) > 2 else 384
H, margin = 384, 0.3


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def twin(m, dtype):
    ref = O.OracleTwoTowers(P, train_table=True)
    with torch.no_grad():
        for t_new, t_ref in ((m.query_tower, ref.query_tower), (m.document_tower, ref.document_tower)):
            t_ref.table.copy_(t_new.pretrained_model.table.detach().float().cpu())
            for i in (0, 2):
                t_ref.projection[i].weight.copy_(t_new.projection[i].weight.cpu())
                t_ref.projection[i].bias.copy_(t_new.projection[i].bias.cpu())
    return ref.to(dtype)


def backward(ref, batch, gate):
    for prm in ref.parameters():
        prm.grad = None
    ys, xs = [], []
    for g, (tower, ids, mask) in enumerate(((ref.query_tower, batch.q_ids, batch.q_mask),
                                            (ref.document_tower, batch.p_ids, batch.p_mask),
                                            (ref.document_tower, batch.n_ids, batch.n_mask))):
        emb = torch.nn.functional.embedding(ids.long(), tower.table)
        m = mask.long().unsqueeze(-1).expand(emb.size()).to(emb.dtype)
        pooled = torch.sum(emb * m, 1) / torch.clamp(m.sum(1), min=1e-9)
        x = torch.nn.functional.normalize(pooled, p=2, dim=1)
        x.retain_grad()
        z = torch.nn.functional.linear(x, tower.projection[0].weight, tower.projection[0].bias)
        h = z * gate[g * B: (g + 1) * B].to(z.dtype)
        ys.append(torch.nn.functional.linear(h, tower.projection[2].weight, tower.projection[2].bias))
        xs.append(x)
    loss = O.triplet_loss(ys[0], ys[1], ys[2], margin)
    loss.backward()
    return torch.cat([x.grad for x in xs]), (ref.query_tower.table.grad, ref.document_tower.table.grad)


batch = O.synth_triplet_batch(B, 32, 256, "Z", seed=12)
counts = (torch.bincount(batch.q_ids[batch.q_mask > 0].flatten().long(), minlength=O.VOCAB),
          torch.bincount(torch.cat([batch.p_ids[batch.p_mask > 0], batch.n_ids[batch.n_mask > 0]]).flatten().long(),
                         minlength=O.VOCAB))
for precision in sys.argv[3:] or ["fp32", "bf16x3"]:
    torch.manual_seed(2)
    m = TwoTowersModel(projection_dim=P, precision=precision, train_table=True).cuda()
    Lq, Ld = batch.q_ids.shape[1], batch.p_ids.shape[1]
    tr = FusedTrainer(m, margin, 1e-3, B, Lq, Ld, precision=precision, use_graph=False, ids_dtype=torch.int32,
                      mask_dtype=torch.uint8)
    for dst, src in zip(tr.tok, batch.astuple()):
        dst.copy_(src.to(dst.dtype))
    tr._fwd_bwd()
    torch.cuda.synchronize()
    gate = tr.step_obj.relu_gate().cpu()
    dx64, tg64 = backward(twin(m, torch.float64), batch, gate)
    dx32, tg32 = backward(twin(m, torch.float32), batch, gate)
    v = tr.step_obj._views()
    print(f"--- {precision} B={B} P={P}")
    if "dxhat" in v:
        for g, nm in enumerate("qpn"):
            print(f"  dxhat {nm}: device {rel(v['dxhat'][g * B:(g + 1) * B], dx64[g * B:(g + 1) * B]):.2e}   "
                  f"fp32 oracle {rel(dx32[g * B:(g + 1) * B], dx64[g * B:(g + 1) * B]):.2e}")
    for t, nm in enumerate(("query", "document")):
        heavy = counts[t] > 256
        got = tr.table_grads[t].cpu()
        for tag, sel in (("all", slice(None)), ("heavy", heavy), ("light", ~heavy)):
            print(f"  table {nm:8s} {tag:5s}: device {rel(got[sel], tg64[t][sel]):.2e}   fp32 oracle "
                  f"{rel(tg32[t][sel], tg64[t][sel]):.2e}   device vs fp32 oracle {rel(got[sel], tg32[t][sel]):.2e}   "
                  f"|g| {float(tg64[t][sel].norm()):.3e}   rows {int(heavy.sum()) if tag == 'heavy' else ''}")
        worst = ((got.double() - tg64[t]).norm(dim=1) / tg64[t].norm(dim=1).clamp_min(1e-30))
        worst[tg64[t].norm(dim=1) == 0] = 0
        top = torch.topk(worst, 5)
        print("   worst rows:", [(int(i), int(counts[t][i]), f"{float(e):.1e}") for e, i in zip(top.values, top.indices)])
    del tr, m
