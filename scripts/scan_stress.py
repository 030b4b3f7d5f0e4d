"""Stress: rank `r` of `w` shard of bench.py's synthetic corpus, many streaming (128-query) searches."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from two_towers_overlords_b200 import retrieval

w, r, iters = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
n_docs, n_queries, P = 8_800_000, 100_000, 384
dev = torch.device("cuda")
lo, hi = retrieval.shard_bounds(n_docs, w, r)
g = torch.Generator(device=dev).manual_seed(7)
gq = torch.Generator(device=dev).manual_seed(8)
Qe = torch.randn(n_queries, P, generator=gq, device=dev)
gr = torch.Generator(device=dev).manual_seed(9)
n_rel = torch.randint(1, 11, (n_queries,), generator=gr, device=dev)
owner = torch.repeat_interleave(torch.arange(n_queries, device=dev), n_rel)
rel_ids = torch.randperm(n_docs, generator=gr, device=dev)[: owner.numel()]
noise = 1.0 + 5.0 * torch.rand(owner.numel(), generator=gr, device=dev)
De = torch.empty(hi - lo, P, device=dev)
chunk = 1 << 20
for s0 in range(0, n_docs, chunk):
    blk = torch.randn(min(chunk, n_docs - s0), P, generator=g, device=dev)
    a, b = max(lo, s0), min(hi, s0 + blk.shape[0])
    if a < b:
        De[a - lo: b - lo] = blk[a - s0: b - s0]
mine = (rel_ids >= lo) & (rel_ids < hi)
gn = torch.Generator(device=dev).manual_seed(10)
plant_noise = torch.randn(owner.numel(), P, generator=gn, device=dev)
De[rel_ids[mine] - lo] = Qe[owner[mine]] + noise[mine, None] * plant_noise[mine]
del plant_noise
shard = retrieval.CorpusShard(De, id_base=lo, precision="bf16")
del De
torch.cuda.synchronize()
q_small = Qe[:128].contiguous()
ref = None
t0 = time.time()
for it in range(iters):
    s, i = shard.search(q_small, 10)
    if it % 50 == 0:
        torch.cuda.synchronize()
        if ref is None: ref = i.clone()
        assert torch.equal(i, ref), f"iteration {it}: ids changed"
torch.cuda.synchronize()
print(f"rank {r}/{w}: {iters} streaming searches ok in {time.time() - t0:.1f} s", flush=True)
s, i = shard.search(Qe, 10)
torch.cuda.synchronize()
print("batched ok", flush=True)
