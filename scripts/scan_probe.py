"""GPU probe: one corpus-scan pass at a given size (for ncu launch lists and quick timing)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from two_towers_overlords_b200 import ops, retrieval

N, Q, P = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 384
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
prec = sys.argv[5] if len(sys.argv) > 5 else "bf16"
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
De = torch.randn(N, P, generator=g, device=dev)
Qe = torch.randn(Q, P, generator=g, device=dev)
shard = retrieval.CorpusShard(De, precision=prec)
del De
for _ in range(2):
    shard.search(Qe, 10)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    s, i = shard.search(Qe, 10)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"N={N} Q={Q} P={P} {prec}: {ms:.3f} ms/pass, {Q/ms*1e3:.0f} q/s, {2*Q*N*P/ms/1e9:.1f} TFLOP/s, stream {N*P*2/ms/1e6:.1f} GB/s")
