"""Timing probe (results are NOT numerically meaningful: the two branches race on the activation buffers):
how long does a step take when the pooled gather of the next step runs beside the rest of the current one?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from two_towers_overlords_b200 import TwoTowersModel
from two_towers_overlords_b200.training import FusedTrainer

dev = torch.device("cuda")
torch.manual_seed(0)
B, P = 2048, 512
model = TwoTowersModel(projection_dim=P, precision="bf16x3").to(dev)
tr = FusedTrainer(model, 0.3, 1e-3, B, 32, 256, precision="bf16x3", use_graph=False, token_slots=8)
for slot in range(8):
    for t in tr.tok_slots[slot]:
        if t.dtype == torch.uint8: t.fill_(1)
        else: t.copy_(torch.randint(999, 30522, t.shape, device=dev).to(t.dtype))
side = torch.cuda.Stream(priority=int(sys.argv[1]) if len(sys.argv) > 1 else 0)
main = torch.cuda.current_stream()

def seq(slot):
    tr._fwd_bwd(slot, 1); tr._fwd_bwd(slot, 2); tr._optimizer()
def ovl(slot):
    side.wait_stream(main)
    with torch.cuda.stream(side):
        tr._fwd_bwd((slot + 1) % 8, 1)
    tr._fwd_bwd(slot, 2); tr._optimizer()
    main.wait_stream(side)
def only_back(slot):
    tr._fwd_bwd(slot, 2); tr._optimizer()
def only_front(slot):
    tr._fwd_bwd(slot, 1)

for name, fn in (("sequential", seq), ("overlapped", ovl), ("back only", only_back), ("front only", only_front)):
    for i in range(3): fn(i % 8)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(8): fn(i)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 160 * 1e3:.1f} us/step")
