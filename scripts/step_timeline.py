"""Kernel timeline of the pipelined training step (torch.profiler / CUPTI): start offset, duration and stream of every
kernel of one steady-state step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from two_towers_overlords_b200 import TwoTowersModel
from two_towers_overlords_b200.training import FusedTrainer

dev = torch.device("cuda")
torch.manual_seed(0)
B, P, NS = 2048, 512, 8
model = TwoTowersModel(projection_dim=P, precision="bf16x3").to(dev)
tr = FusedTrainer(model, 0.3, 1e-3, B, 32, 256, precision="bf16x3", token_slots=NS, ids_dtype=torch.uint16)
for slot in range(NS):
    for t in tr.tok_slots[slot]:
        if t.dtype == torch.uint8: t.fill_(1)
        else: t.copy_(torch.randint(999, 30522, t.shape, device=dev).to(t.dtype))
tr.prepare()
for i in range(24): tr.step(i % NS, (i + 1) % NS)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(24, 32): tr.step(i % NS, (i + 1) % NS)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# steady state: from the end of the 4th step-closing kernel (Adam launch, or the chain kernel that carries Adam) to the
# end of the 5th
closers = [i for i, e in enumerate(evs) if "adam_dev" in e.name] or [i for i, e in enumerate(evs) if "chain_kernel" in e.name]
a0, a1 = closers[3], closers[4]
t0 = evs[a0].time_range.end
print(f"step = [end of step-closing kernel #3, end of #4] = {evs[a1].time_range.end - t0:.1f} us")
for e in evs[a0 + 1: a1 + 1]:
    name = e.name.split("(")[0].split("::")[-1][:34]
    print(f"{e.time_range.start - t0:8.1f} +{e.time_range.end - e.time_range.start:7.1f} us  {name}")
