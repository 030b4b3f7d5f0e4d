"""
Exhaustive corpus-scan retrieval over a (sharded) document matrix, feeding NDCG@k.

The reference evaluates one query at a time on the CPU (backend/training.py:244-311) and serves through an
approximate Redis HNSW index (backend/search.py:352-402).  Here the whole query batch is scored against
the corpus shard resident in this GPU's HBM by tt_scan_topk (score matrix never stored), and with
world_size > 1 the per-shard top-k lists are exchanged with ONE all-gather and merged by tt_topk_merge
(order: score descending, id ascending — identical for any shard count).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

try:
    from . import ops
except ImportError:
    import ops


def shard_bounds(n_docs: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of rank; the last rank takes the remainder."""
    per = n_docs // world_size
    lo = rank * per
    hi = n_docs if rank == world_size - 1 else lo + per
    return lo, hi


class CorpusShard:
    """Normalised document embeddings of one shard, resident on the device (fp32 + optional bf16 copy)."""

    def __init__(self, doc_embeds: torch.Tensor, id_base: int = 0, precision: str = "fp32"):
        self.precision = precision
        self.id_base = int(id_base)
        if precision == "fp32":
            self.Dn, self.Db = ops.l2_normalize_rows(doc_embeds), None
        else:
            self.Dn, self.Db = ops.l2_normalize_rows(doc_embeds, want_bf16=True)

    def __len__(self):
        return self.Dn.shape[0]

    def search(self, query_embeds: torch.Tensor, k: int = 10):
        """-> (scores [Q,k], global ids [Q,k]) of this shard only."""
        if self.precision == "fp32":
            Qn, Qb = ops.l2_normalize_rows(query_embeds), None
        else:
            Qn, Qb = ops.l2_normalize_rows(query_embeds, want_bf16=True)
        return ops.scan_topk(Qn, self.Dn, k=k, id_base=self.id_base, precision=self.precision, Qb=Qb, Db=self.Db)


def all_gather_lists(top_s: torch.Tensor, top_i: torch.Tensor, world_size: int, group=None):
    """The one exchange of the sharded scan: per-shard lists [Q,k] -> [G,Q,k] on every rank."""
    import torch.distributed as dist

    parts_s = torch.empty((world_size,) + tuple(top_s.shape), dtype=top_s.dtype, device=top_s.device)
    parts_i = torch.empty((world_size,) + tuple(top_i.shape), dtype=top_i.dtype, device=top_i.device)
    dist.all_gather(list(parts_s.unbind(0)), top_s.contiguous(), group=group)
    dist.all_gather(list(parts_i.unbind(0)), top_i.contiguous(), group=group)
    return parts_s, parts_i


def gather_and_merge(top_s: torch.Tensor, top_i: torch.Tensor, world_size: int, group=None, peer=None):
    """All-gather the per-shard lists [Q,k] and merge them to the global top-k on every rank.  `peer`
    (comm.PeerLists) merges them in place over NVLink peer memory instead of the NCCL all-gather."""
    if world_size == 1:
        return top_s, top_i
    if peer is not None:
        return peer.merge(top_s, top_i)
    return ops.topk_merge(*all_gather_lists(top_s, top_i, world_size, group))


def relevance_csr(relevant: Sequence[Sequence[int]], device) -> tuple[torch.Tensor, torch.Tensor]:
    """Relevant-document sets as CSR (offsets [Q+1], sorted ids) for tt_ndcg_at_k."""
    lens = np.fromiter((len(r) for r in relevant), dtype=np.int64, count=len(relevant))
    offs = np.concatenate([[0], np.cumsum(lens)])
    flat = np.fromiter((d for r in relevant for d in sorted(r)), dtype=np.int64, count=int(offs[-1]))
    return torch.from_numpy(offs).to(device), torch.from_numpy(flat).to(device)


def retrieve_and_score(shard: CorpusShard, query_embeds: torch.Tensor, relevant_csr, k: int = 10,
                       world_size: int = 1, group=None, peer=None):
    """Top-k ids of every query over the whole (sharded) corpus and their NDCG@k."""
    top_s, top_i = shard.search(query_embeds, k)
    top_s, top_i = gather_and_merge(top_s, top_i, world_size, group, peer)
    ndcg = ops.ndcg_at_k(top_i, relevant_csr[0], relevant_csr[1], kk=k)
    return top_s, top_i, ndcg
