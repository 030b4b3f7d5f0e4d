"""Timing probe: one 2048-triplet step vs two concurrent 1024-triplet half-steps (separate streams, one graph)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from two_towers_overlords_b200 import TwoTowersModel
from two_towers_overlords_b200.training import FusedTrainer

dev = torch.device("cuda")
torch.manual_seed(0)
P, NSLOT = 512, 8
def make(B):
    model = TwoTowersModel(projection_dim=P, precision="bf16x3").to(dev)
    tr = FusedTrainer(model, 0.3, 1e-3, B, 32, 256, precision="bf16x3", use_graph=False, token_slots=NSLOT)
    for slot in range(NSLOT):
        for t in tr.tok_slots[slot]:
            if t.dtype == torch.uint8: t.fill_(1)
            else: t.copy_(torch.randint(999, 30522, t.shape, device=dev).to(t.dtype))
    tr._warm_up()
    return tr
def timeit(name, g, per, n=20):
    for i in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / n / per * 1e3:.1f} us per 2048 triplets", flush=True)

cap = torch.cuda.Stream(priority=-1)
full = make(2048)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=cap):
    for s in range(NSLOT): full._pipelined(s, (s + 1) % NSLOT, s & 1)
timeit("one 2048 chain", g, NSLOT)
for k in (2, 4):
    halves = [make(2048 // k) for _ in range(k)]
    streams = [torch.cuda.Stream(priority=-1) for _ in range(k)]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=cap):
        cur = torch.cuda.current_stream()
        for s in range(NSLOT):
            for st in streams: st.wait_stream(cur)
            for h, st in zip(halves, streams):
                with torch.cuda.stream(st):
                    h._pipelined(s, (s + 1) % NSLOT, s & 1)
            for st in streams: cur.wait_stream(st)
    timeit(f"{k} concurrent {2048 // k}-triplet chains", g, NSLOT)
