"""Multi-process check of the peer-memory exchange (run under torchrun on an NVLink box):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dp_peer_check.py
Trains the same model for a few steps with exchange="peer" (fused kernel over CUDA IPC) and exchange="nccl"
(all-reduce + Adam) on identical data and compares parameters and losses; also checks replicas stay bit-identical."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)



def synth_batch(B, Lq, Ld, seed, vocab):
    """Shape-U synthetic triplet tokens: (q_ids, q_mask, p_ids, p_mask, n_ids, n_mask), negatives = rolled positives."""
    g = torch.Generator().manual_seed(seed)
    q = torch.randint(999, vocab, (B, Lq), generator=g, dtype=torch.int64)
    p = torch.randint(999, vocab, (B, Ld), generator=g, dtype=torch.int64)
    n = torch.roll(p, 1 + seed % (B - 1), 0)
    return tuple(t for ids in (q, p, n) for t in (ids, torch.ones_like(ids)))


from two_towers_overlords_b200 import TwoTowersModel  # noqa: E402
from two_towers_overlords_b200.training import FusedTrainer  # noqa: E402

B, Lq, Ld, P, V, STEPS = 512, 32, 128, 256, 30522, 6
res = {}
for exchange in ("nccl", "peer", "peer-pipelined"):
    torch.manual_seed(0)
    model = TwoTowersModel(projection_dim=P, vocab_size=V, precision="bf16x3").to(dev)
    tr = FusedTrainer(model, 0.3, 1e-3, B, Lq, Ld, precision="bf16x3", world_size=world, rank=rank,
                      ids_dtype=torch.int64, mask_dtype=torch.int64, exchange=exchange.split("-")[0], token_slots=2)
    losses = []
    batches = [synth_batch(B, Lq, Ld, 100 * i + rank, V) for i in range(STEPS)]
    tr.load_packed(tr.pack_host_tokens(batches[0], pin=False), 0)
    for i in range(STEPS):
        nslot = None
        if i + 1 < STEPS:
            if exchange == "peer-pipelined":
                nslot = (i + 1) % 2
            tr.load_packed(tr.pack_host_tokens(batches[i + 1], pin=False), (i + 1) % 2)
        tr.step(i % 2, nslot)
        losses.append(float(tr.loss_view[0].item()))
    torch.cuda.synchronize()
    p = tr.flat_p.clone()
    # replicas identical?
    gathered = [torch.empty_like(p) for _ in range(world)]
    dist.all_gather(gathered, p)
    same = all(torch.equal(g, gathered[0]) for g in gathered)
    res[exchange] = (p, losses, same)
    dist.barrier()
    tr.close()
    dist.barrier()

pn, ln, sn = res["nccl"]
pp, lp, sp = res["peer"]
err = float((pn - pp).norm() / pn.norm())
if rank == 0:
    print(f"world={world} replicas identical: nccl={sn} peer={sp}; |p_peer - p_nccl|/|p| = {err:.3e}")
    print("loss nccl:", [f"{x:.6f}" for x in ln])
    print("loss peer:", [f"{x:.6f}" for x in lp])
pq, lq, sq = res["peer-pipelined"]
assert sq and torch.equal(pq, pp), "pipelined peer exchange differs from the step-by-step one"
assert sp, "peer exchange left the replicas different"
assert err < 1e-5, err
assert all(abs(a - b) <= 1e-5 * abs(a) for a, b in zip(ln, lp))
# corpus-scan list merge over peer memory vs all-gather + tt_topk_merge
from two_towers_overlords_b200 import comm, retrieval  # noqa: E402

Q, k = 5000, 10
g = torch.Generator(device=dev).manual_seed(77 + rank)
top_s = torch.randn(Q, k, generator=g, device=dev).sort(dim=1, descending=True).values
top_i = (torch.arange(Q * k, device=dev).reshape(Q, k) * world + rank)
peer = comm.PeerLists(Q, k, world, rank, dev)
for rep in range(3):
    a_s, a_i = retrieval.gather_and_merge(top_s + rep, top_i, world)
    b_s, b_i = retrieval.gather_and_merge(top_s + rep, top_i, world, peer=peer)
    torch.cuda.synchronize()
    peer.check()
    assert torch.equal(a_i, b_i) and torch.equal(a_s, b_s), "peer list merge differs from all-gather + merge"
dist.barrier()
peer.close()
if rank == 0:
    print("peer list merge == all-gather + merge")
    print("OK")
dist.destroy_process_group()
