"""Diagnostic: where does the gather/rest overlap get lost? (timing only)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from two_towers_overlords_b200 import TwoTowersModel
from two_towers_overlords_b200.training import FusedTrainer

dev = torch.device("cuda")
torch.manual_seed(0)
B, P, NSLOT = 2048, 512, 8
def make(use_graph):
    model = TwoTowersModel(projection_dim=P, precision="bf16x3").to(dev)
    tr = FusedTrainer(model, 0.3, 1e-3, B, 32, 256, precision="bf16x3", use_graph=use_graph, token_slots=NSLOT)
    for slot in range(NSLOT):
        for t in tr.tok_slots[slot]:
            if t.dtype == torch.uint8: t.fill_(1)
            else: t.copy_(torch.randint(999, 30522, t.shape, device=dev).to(t.dtype))
    return tr
def timeit(name, fn, n=160):
    for i in range(16): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / n * 1e3:.1f} us/step", flush=True)

tr = make(True); tr.prepare()
timeit("trainer graph, step(s, s+1)", lambda i: tr.step(i % NSLOT, (i + 1) % NSLOT))
timeit("trainer graph, step(s, None)", lambda i: tr.step(i % NSLOT, None))
tr2 = make(False)
timeit("trainer eager, step(s, s+1)", lambda i: tr2.step(i % NSLOT, (i + 1) % NSLOT))
# manual graphs from the trainer's own _pipelined
gs = []
for s in range(NSLOT):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        tr2._pipelined(s, (s + 1) % NSLOT, s & 1)
    gs.append(g)
timeit("8 separate graphs of _pipelined", lambda i: gs[i % NSLOT].replay())
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for s in range(NSLOT): tr2._pipelined(s, (s + 1) % NSLOT, s & 1)
timeit("one graph of 8 x _pipelined", lambda i: g.replay(), n=20)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for s in range(NSLOT): tr2._pipelined(s, (s + 1) % NSLOT, 0)
timeit("one graph of 8 x _pipelined, parity 0 only (front writes ws 1, back reads ws 0)", lambda i: g.replay(), n=20)
