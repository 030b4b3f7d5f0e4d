"""Diagnostic: gather/rest overlap with the rest of the step on a high-priority stream (timing only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from two_towers_overlords_b200 import TwoTowersModel
from two_towers_overlords_b200.training import FusedTrainer

dev = torch.device("cuda")
torch.manual_seed(0)
B, P, NSLOT = 2048, 512, 8
model = TwoTowersModel(projection_dim=P, precision="bf16x3").to(dev)
tr = FusedTrainer(model, 0.3, 1e-3, B, 32, 256, precision="bf16x3", use_graph=False, token_slots=NSLOT)
for slot in range(NSLOT):
    for t in tr.tok_slots[slot]:
        if t.dtype == torch.uint8: t.fill_(1)
        else: t.copy_(torch.randint(999, 30522, t.shape, device=dev).to(t.dtype))
tr._warm_up()
def timeit(name, fn, n=20):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / n / NSLOT * 1e3:.1f} us/step", flush=True)

for cap_prio, side_prio in ((0, 0), (-1, 0), (0, -1)):
    cap = torch.cuda.Stream(priority=cap_prio)
    tr.side = torch.cuda.Stream(priority=side_prio)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=cap):
        for s in range(NSLOT): tr._pipelined(s, (s + 1) % NSLOT, s & 1)
    timeit(f"capture prio {cap_prio}, gather prio {side_prio}", lambda i: g.replay())
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for s in range(NSLOT): tr._pipelined(s, None, 0)
timeit("rest only", lambda i: g.replay())
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for s in range(NSLOT): tr._fwd_bwd(s, 1, 0)
timeit("gather only", lambda i: g.replay())
