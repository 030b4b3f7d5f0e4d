"""Per-phase timeline of the persistent chain kernel (TT_CHAIN_TRACE=1): for every task phase, when its first task
started and its last task ended (us from the kernel's first task), and the mean / max epilogue-side task duration.
GPU box only:  python scripts/chain_timeline.py [B] [P] [precision] [table]"""
import ctypes
import os
import sys

import numpy as np
import torch

os.environ["TT_CHAIN_TRACE"] = "1"
os.environ["TT_CHAIN"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from two_towers_overlords_b200 import TwoTowersModel, _native as N  # noqa: E402
from two_towers_overlords_b200.training import FusedTrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
P = int(sys.argv[2]) if len(sys.argv) > 2 else 512
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16x3"
table = len(sys.argv) > 4 and sys.argv[4] == "table"
H, Lq, Ld = 384, 32, 256
torch.manual_seed(0)
m = TwoTowersModel(projection_dim=P, precision=prec, train_table=table).cuda()
tr = FusedTrainer(m, 0.3, 1e-3, B, Lq, Ld, precision=prec, use_graph=False, ids_dtype=torch.int32,
                  mask_dtype=torch.uint8)
g = torch.Generator().manual_seed(1)
for t in tr.tok:
    if t.dtype == torch.uint8:
        t.fill_(1)
    else:
        t.copy_(torch.randint(999, 30522, t.shape, generator=g).to(t.dtype))
for _ in range(3):
    tr._fwd_bwd()
torch.cuda.synchronize()
tr._fwd_bwd(0, 2, 0)
torch.cuda.synchronize()
so = tr.step_obj
n_ctas, slots = 160, 64
raw = so.internal("trace", n_ctas * slots, 2, dtype=torch.int64).cpu().numpy().reshape(n_ctas, slots, 2)
RTB, NC, NCH = (B + 127) // 128, (P + 127) // 128, (H + 127) // 128
sms = torch.cuda.get_device_properties(0).multi_processor_count
kbt = [(B + 63) // 64, (2 * B + 63) // 64]
kcb = 8
while sum((k + kcb - 1) // kcb for k in kbt) > 48:
    kcb += 8
nch = sum((k + kcb - 1) // kcb for k in kbt)
counts = [sms, 0, 3 * RTB * NC, 3 * RTB * (P // 64), 3 * RTB * NC, 3 * RTB * NCH if table else 0, nch * NC * NC, nch * NC * NCH, sms]
names = ["S", "T", "F1", "F2L", "DZ", "DX", "DW2", "DW1", "G"]
offs = np.concatenate([[0], np.cumsum(counts)])
t0 = raw[:, :, 1][raw[:, :, 1] > 0].min()
print(f"B={B} P={P} {prec} table={table}: {offs[-1]} tasks on {sms} CTAs, kcb={kcb}")
ev = {}
for c in range(n_ctas):
    for s in range(slots):
        tag, t = int(raw[c, s, 0]), int(raw[c, s, 1])
        if t == 0:
            continue
        ev.setdefault(tag >> 2, {})[("start", "end", "acc", "sib")[tag & 3]] = (t - t0) / 1e3
for k, nm in enumerate(names):
    rows = [ev[i] for i in range(offs[k], offs[k + 1]) if i in ev and "start" in ev[i] and "end" in ev[i]]
    if not rows:
        continue
    st = np.array([r["start"] for r in rows]); en = np.array([r["end"] for r in rows])
    extra = ""
    if all("acc" in r for r in rows):
        ac = np.array([r["acc"] for r in rows])
        extra = f"  wait-for-acc mean {np.mean(ac - st):6.1f}  epilogue mean {np.mean(en - ac):6.1f} max {np.max(en - ac):6.1f}"
    if all("sib" in r for r in rows):
        sb = np.array([r["sib"] for r in rows])
        extra += f"  pass1 {np.mean(sb - ac):5.1f} (incl. sibling wait)  pass2 {np.mean(en - sb):5.1f}"
    print(f"  {nm:4s} {len(rows):4d} tasks: first start {st.min():7.1f}  last start {st.max():7.1f}  first end {en.min():7.1f}  "
          f"last end {en.max():7.1f}  task us mean {np.mean(en - st):6.1f} max {np.max(en - st):6.1f}{extra}")
