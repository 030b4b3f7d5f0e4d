"""
CPU oracle for the two-tower hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import
this module, and only as the checker or the timed CPU baseline.  Nothing under
two-towers-overlords_b200/ imports it; the product path fails loudly when libtt_b200.so is missing.

What it restates (fp32 torch-CPU ops, the same ATen arithmetic the reference runs on CPU):
  * backend/model.py:40-72     AveragePoolingTower.forward with the north_star backbone
                               `pretrained_model(**tokens)[0] == E[input_ids]`
  * backend/model.py:125-145   TripletLoss (cosine distance, margin hinge, batch mean)
  * backend/training.py:36-53  one fp32 training step (zero_grad, 3 encodes, loss, backward, Adam)
  * backend/training.py:297-311 per-query cosine scoring + sklearn ndcg_score(k)
  * sklearn/metrics/_ranking.py (scikit-learn, reference pin 1.7.0 in backend/uv.lock:1337, 1.9.0
    installed): _dcg_sample_scores / _tie_averaged_dcg / _ndcg_sample_scores, restated in numpy
    because scikit-learn is a third-party dependency outside /root/reference.

Parity pin: the reference ships no tests or golden vectors for this path, so the
oracle is pinned against outputs of the reference's OWN modules imported from /root/reference
(oracle/gen_golden.py, run in the build container; vectors committed under tests/golden/).
tests/test_oracle_golden.py checks every function here against those vectors.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F

HIDDEN = 384
VOCAB = 30522


# ----------------------------------------------------------------------------------------------
# tower forward
# ----------------------------------------------------------------------------------------------
def mean_pooling(token_embeddings: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """backend/model.py:63-72 — sum(h * mask) / clamp(sum(mask), 1e-9), mask broadcast over H."""
    m = attention_mask.unsqueeze(-1).expand(token_embeddings.size()).float()
    return torch.sum(token_embeddings * m, 1) / torch.clamp(m.sum(1), min=1e-9)


def pooled_normalised(table: torch.Tensor, ids: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """backend/model.py:51-56 with the embedding-only backbone: E[ids] -> masked mean -> L2 normalise."""
    h = F.embedding(ids.long(), table.float())
    return F.normalize(mean_pooling(h, mask.long()), p=2, dim=1)


def projection(x: torch.Tensor, W1, b1, W2, b2) -> torch.Tensor:
    """backend/model.py:33-38,59 — Linear(H,P) -> ReLU -> Linear(P,P)."""
    return F.linear(F.relu(F.linear(x, W1, b1)), W2, b2)


def tower_forward(table, W1, b1, W2, b2, ids, mask) -> torch.Tensor:
    """backend/model.py:40-60."""
    return projection(pooled_normalised(table, ids, mask), W1, b1, W2, b2)


# ----------------------------------------------------------------------------------------------
# loss
# ----------------------------------------------------------------------------------------------
def cosine_distance(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """backend/model.py:132-135."""
    return 1 - F.cosine_similarity(x, y, dim=1)


def triplet_loss(anchor, positive, negative, margin: float) -> torch.Tensor:
    """backend/model.py:137-145."""
    return F.relu(cosine_distance(anchor, positive) - cosine_distance(anchor, negative) + margin).mean()


# ----------------------------------------------------------------------------------------------
# model container mirroring the reference's parameter names
# ----------------------------------------------------------------------------------------------
class OracleTower(torch.nn.Module):
    """AveragePoolingTower (backend/model.py:13-60) over pre-tokenised input."""

    def __init__(self, projection_dim: int, vocab: int = VOCAB, hidden: int = HIDDEN,
                 train_table: bool = False):
        super().__init__()
        self.table = torch.nn.Parameter(torch.empty(vocab, hidden).normal_(), requires_grad=train_table)
        self.projection = torch.nn.Sequential(
            torch.nn.Linear(hidden, projection_dim), torch.nn.ReLU(),
            torch.nn.Linear(projection_dim, projection_dim))

    def forward(self, ids: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        if self.table.requires_grad:  # D2 extension: the reference's no_grad (model.py:51) lifted
            h = F.embedding(ids.long(), self.table)
        else:
            with torch.no_grad():
                h = F.embedding(ids.long(), self.table)
        x = F.normalize(mean_pooling(h, mask.long()), p=2, dim=1)
        return self.projection(x)


class OracleTwoTowers(torch.nn.Module):
    """TwoTowersModel (backend/model.py:75-121): two independent towers, own table copies."""

    def __init__(self, projection_dim: int, vocab: int = VOCAB, hidden: int = HIDDEN,
                 train_table: bool = False):
        super().__init__()
        self.query_tower = OracleTower(projection_dim, vocab, hidden, train_table)
        self.document_tower = OracleTower(projection_dim, vocab, hidden, train_table)

    def encode_queries(self, ids, mask):
        return self.query_tower(ids, mask)

    def encode_documents(self, ids, mask):
        return self.document_tower(ids, mask)


def train_step(model: OracleTwoTowers, optimizer: torch.optim.Optimizer, batch, margin: float) -> float:
    """backend/training.py:36-53 for one batch; batch = (q_ids,q_mask,p_ids,p_mask,n_ids,n_mask)."""
    q_ids, q_mask, p_ids, p_mask, n_ids, n_mask = batch
    optimizer.zero_grad()
    q = model.encode_queries(q_ids, q_mask)
    p = model.encode_documents(p_ids, p_mask)
    n = model.encode_documents(n_ids, n_mask)
    loss = triplet_loss(q, p, n, margin)
    loss.backward()
    optimizer.step()
    return loss.item()


def train_epoch_accum(model: OracleTwoTowers, optimizer: torch.optim.Optimizer, batches, margin: float,
                      accumulation_steps: int = 2) -> float:
    """backend/training.py:81-133 (train_epoch_optimized) in fp32: loss / accumulation_steps, gradients accumulate
    across batches, optimiser step + zero_grad every `accumulation_steps` batches (a trailing partial group is
    accumulated but never stepped), returns the mean of the unscaled per-batch losses.  autocast and GradScaler
    are left out: the scaler multiplies by a power of two and divides it back (exact in fp32), and the CUDA path
    this oracle checks computes in fp32 / split-bf16 regardless of autocast."""
    total, nb = 0.0, 0
    for idx, (q_ids, q_mask, p_ids, p_mask, n_ids, n_mask) in enumerate(batches):
        q = model.encode_queries(q_ids, q_mask)
        p = model.encode_documents(p_ids, p_mask)
        n = model.encode_documents(n_ids, n_mask)
        loss = triplet_loss(q, p, n, margin) / accumulation_steps
        loss.backward()
        if (idx + 1) % accumulation_steps == 0:
            optimizer.step()
            optimizer.zero_grad()
        total += loss.item() * accumulation_steps
        nb += 1
    return total / nb


# ----------------------------------------------------------------------------------------------
# evaluation
# ----------------------------------------------------------------------------------------------
def cosine_scores(query_embed: torch.Tensor, doc_embeds: torch.Tensor) -> np.ndarray:
    """backend/training.py:297-299 — one query [1,P] against all docs [N,P]."""
    return torch.cosine_similarity(query_embed.cpu(), doc_embeds.cpu(), dim=1).numpy()


def _dcg_sample(y_true: np.ndarray, y_score: np.ndarray, k: int | None, ignore_ties: bool) -> float:
    """sklearn _dcg_sample_scores for one sample (log base 2)."""
    n = y_true.shape[0]
    discount = 1.0 / (np.log(np.arange(n) + 2) / np.log(2))
    if k is not None:
        discount[k:] = 0
    if ignore_ties:
        ranking = np.argsort(y_score)[::-1]
        return float(discount.dot(y_true[ranking]))
    # _tie_averaged_dcg
    discount_cumsum = np.cumsum(discount)
    _, inv, counts = np.unique(-y_score, return_inverse=True, return_counts=True)
    ranked = np.zeros(len(counts))
    np.add.at(ranked, inv, y_true)
    ranked /= counts
    groups = np.cumsum(counts) - 1
    discount_sums = np.empty(len(counts))
    discount_sums[0] = discount_cumsum[groups[0]]
    discount_sums[1:] = np.diff(discount_cumsum[groups])
    return float((ranked * discount_sums).sum())


def ndcg_at_k(y_true: np.ndarray, y_score: np.ndarray, k: int | None = None) -> float:
    """sklearn.metrics.ndcg_score for ONE sample as called at backend/training.py:304-309
    (default ignore_ties=False): tie-averaged DCG / ideal DCG (ties ignored), 0 when ideal is 0."""
    y_true = np.asarray(y_true, dtype=np.float64).ravel()
    y_score = np.asarray(y_score).ravel()
    if y_true.shape[0] <= 1:
        raise ValueError("Computing NDCG is only meaningful when there is more than 1 document.")
    if y_true.min() < 0:
        raise ValueError("ndcg_score should not be used on negative y_true values.")
    gain = _dcg_sample(y_true, y_score, k, ignore_ties=False)
    ideal = _dcg_sample(y_true, y_true, k, ignore_ties=True)
    return 0.0 if ideal == 0 else gain / ideal


def topk_ids(scores: np.ndarray, k: int) -> np.ndarray:
    """The k best document indices by (score descending, id ascending) — the tie-break convention
    the CUDA scan uses; independent of it whenever scores are distinct."""
    order = np.lexsort((np.arange(scores.shape[0]), -scores.astype(np.float64)))
    return order[:k]


def ndcg_from_topk(top_ids: np.ndarray, relevant: set, k: int) -> float:
    """Closed form of ndcg_at_k for binary relevance and tie-free scores (SURVEY.md §7 hard parts)."""
    dcg = sum(1.0 / math.log2(r + 2) for r, d in enumerate(top_ids[:k]) if int(d) in relevant)
    idcg = sum(1.0 / math.log2(r + 2) for r in range(min(len(relevant), k)))
    return 0.0 if idcg == 0 else dcg / idcg


def retrieval_eval(Qe: torch.Tensor, De: torch.Tensor, relevant: list[set], k: int = 10):
    """The reference's per-query loop (backend/training.py:244-311) over pre-encoded docs:
    returns (top-k ids [Q,k], ndcg [Q])."""
    ids = np.zeros((Qe.shape[0], k), dtype=np.int64)
    ndcg = np.zeros(Qe.shape[0])
    N = De.shape[0]
    for i in range(Qe.shape[0]):
        s = cosine_scores(Qe[i:i + 1], De)
        rel = np.zeros(N)
        rel[list(relevant[i])] = 1
        ndcg[i] = ndcg_at_k(rel, s, k)
        ids[i] = topk_ids(s, k)
    return ids, ndcg


# ----------------------------------------------------------------------------------------------
# synthetic MS MARCO-shaped token batches — shared by tests and bench so both
# sides see identical inputs.
# ----------------------------------------------------------------------------------------------
def synth_tokens(B: int, L: int, shape: str, gen: torch.Generator, kind: str = "doc",
                 vocab: int = VOCAB):
    """shape 'U': all sequences full length, ids uniform in [999,vocab), mask all ones.
    shape 'Z': log-normal lengths, [CLS]=101 ... [SEP]=102, Zipf interior ids, pad id 0 / mask 0."""
    lo = min(999, vocab // 2)
    if shape == "U":
        ids = torch.randint(lo, vocab, (B, L), generator=gen, dtype=torch.int64)
        return ids, torch.ones(B, L, dtype=torch.int64)
    mu, sigma, mn = (math.log(9.0), 0.4, 4) if kind == "query" else (math.log(90.0), 0.45, 16)
    mn = min(mn, L)
    lens = torch.exp(torch.randn(B, generator=gen) * sigma + mu).round().clamp(mn, L).long()
    ranks = torch.arange(1, vocab - lo + 1, dtype=torch.float64)
    probs = 1.0 / ranks
    perm = torch.randperm(vocab - lo, generator=torch.Generator().manual_seed(4242))
    draw = torch.multinomial(probs, B * L, replacement=True, generator=gen).view(B, L)
    ids = perm[draw] + lo
    pos = torch.arange(L).unsqueeze(0)
    mask = (pos < lens.unsqueeze(1)).long()
    ids = ids * mask
    ids[:, 0] = 101
    ids[torch.arange(B), lens - 1] = 102
    Lmax = int(lens.max())
    return ids[:, :Lmax].contiguous(), mask[:, :Lmax].contiguous()


def derangement(B: int, gen: torch.Generator) -> torch.Tensor:
    """In-batch negatives: negative_i = positive_j with j != i (mirrors backend/data.py:113-137)."""
    if B == 1:
        return torch.zeros(1, dtype=torch.long)
    while True:
        perm = torch.randperm(B, generator=gen)
        if not bool((perm == torch.arange(B)).any()):
            return perm


@dataclass
class SynthBatch:
    q_ids: torch.Tensor
    q_mask: torch.Tensor
    p_ids: torch.Tensor
    p_mask: torch.Tensor
    n_ids: torch.Tensor
    n_mask: torch.Tensor
    extra: dict = field(default_factory=dict)

    def astuple(self):
        return (self.q_ids, self.q_mask, self.p_ids, self.p_mask, self.n_ids, self.n_mask)


def synth_triplet_batch(B: int, Lq: int, Ld: int, shape: str, seed: int, vocab: int = VOCAB) -> SynthBatch:
    gen = torch.Generator().manual_seed(seed)
    q_ids, q_mask = synth_tokens(B, Lq, shape, gen, "query", vocab)
    p_ids, p_mask = synth_tokens(B, Ld, shape, gen, "doc", vocab)
    perm = derangement(B, gen)
    return SynthBatch(q_ids, q_mask, p_ids, p_mask, p_ids[perm].contiguous(), p_mask[perm].contiguous())
