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
        self._graphs, self._seen = {}, {}

    def __len__(self):
        return self.Dn.shape[0]

    def _search_eager(self, query_embeds: torch.Tensor, k: int):
        if self.precision == "fp32":
            Qn, Qb = ops.l2_normalize_rows(query_embeds), None
        else:
            Qn, Qb = ops.l2_normalize_rows(query_embeds, want_bf16=True)
        return ops.scan_topk(Qn, self.Dn, k=k, id_base=self.id_base, precision=self.precision, Qb=Qb, Db=self.Db)

    def search(self, query_embeds: torch.Tensor, k: int = 10, graph: bool = True):
        """-> (scores [Q,k], global ids [Q,k]) of this shard only.

        A search is six small launches (normalise, scan, refine, compaction, exact re-scan of unproven queries) plus
        their host-side setup; on a small shard that fixed cost is longer than the scan itself (the latency path
        `DocumentSearchEngine.search` serves).  A query shape seen twice is therefore captured into a CUDA graph —
        one launch per search from then on; results are copies of the graph's static outputs."""
        key = (tuple(query_embeds.shape), int(k), query_embeds.dtype)
        if (not graph or len(self) == 0 or not 0 < query_embeds.shape[0] <= 1024  # big batches are scan-bound anyway
                or torch.cuda.is_current_stream_capturing()):
            return self._search_eager(query_embeds, k)
        ent = self._graphs.get(key)
        if ent is None:
            self._seen[key] = self._seen.get(key, 0) + 1
            if self._seen[key] < 2 or len(self._graphs) >= 8:
                return self._search_eager(query_embeds, k)
            q_static = query_embeds.detach().clone()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out_s, out_i = self._search_eager(q_static, k)
            ent = self._graphs[key] = (g, q_static, out_s, out_i)
        g, q_static, out_s, out_i = ent
        q_static.copy_(query_embeds)
        g.replay()
        return out_s.clone(), out_i.clone()


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
