"""Pins oracle/two_towers_oracle.py against vectors produced by the reference's own modules
(oracle/gen_golden.py; reference: backend/model.py, backend/training.py, sklearn ndcg_score)."""
import os

import numpy as np
import pytest
import torch

from oracle import two_towers_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _oracle_model_from(g, P, prefix=""):
    V = g[prefix + "query_tower__pretrained_model__emb__weight"].shape[0]
    m = O.OracleTwoTowers(P, vocab=V)
    with torch.no_grad():
        for tower in ("query_tower", "document_tower"):
            t = getattr(m, tower)
            t.table.copy_(_t(g[f"{prefix}{tower}__pretrained_model__emb__weight"]))
            for idx in (0, 2):
                t.projection[idx].weight.copy_(_t(g[f"{prefix}{tower}__projection__{idx}__weight"]))
                t.projection[idx].bias.copy_(_t(g[f"{prefix}{tower}__projection__{idx}__bias"]))
    return m


def test_pooling_matches_reference(golden_dir):
    g = _load(golden_dir, "pooling.npz")
    pooled = O.mean_pooling(_t(g["h"]), _t(g["mask"]))
    assert torch.equal(pooled, _t(g["pooled"]))
    normed = torch.nn.functional.normalize(pooled, p=2, dim=1)
    assert torch.equal(normed, _t(g["normed"]))
    assert float(normed[2].abs().max()) == 0.0  # all-masked row stays exactly zero


@pytest.mark.parametrize("tag,P", [("p16", 16), ("p64", 64)])
def test_forward_loss_backward_match_reference(golden_dir, tag, P):
    g = _load(golden_dir, f"step_{tag}.npz")
    m = _oracle_model_from(g, P)
    q = m.encode_queries(_t(g["q_ids"]), _t(g["q_mask"]))
    p = m.encode_documents(_t(g["p_ids"]), _t(g["p_mask"]))
    n = m.encode_documents(_t(g["n_ids"]), _t(g["n_mask"]))
    assert torch.equal(q, _t(g["q"])) and torch.equal(p, _t(g["p"])) and torch.equal(n, _t(g["n"]))
    loss = O.triplet_loss(q, p, n, float(g["margin"]))
    assert loss.item() == pytest.approx(float(g["loss"]), rel=0, abs=0)
    loss.backward()
    for tower in ("query_tower", "document_tower"):
        for idx in (0, 2):
            for kind in ("weight", "bias"):
                ref = _t(g[f"grad__{tower}__projection__{idx}__{kind}"])
                got = getattr(getattr(m, tower).projection[idx], kind).grad
                assert torch.equal(got, ref), (tower, idx, kind)


def test_train_epoch_matches_reference(golden_dir):
    g = _load(golden_dir, "train_epoch.npz")
    m = _oracle_model_from(g, 32, prefix="init__")
    opt = torch.optim.Adam(m.parameters(), lr=float(g["lr"]))
    losses = []
    for b in range(3):
        batch = tuple(_t(g[f"b{b}_{nm}_{k}"]) for nm in ("q", "p", "n") for k in ("ids", "mask"))
        losses.append(O.train_step(m, opt, batch, float(g["margin"])))
    assert np.mean(losses) == pytest.approx(float(g["avg_loss"]), rel=1e-7)
    for tower in ("query_tower", "document_tower"):
        for idx in (0, 2):
            for kind in ("weight", "bias"):
                ref = _t(g[f"final__{tower}__projection__{idx}__{kind}"])
                got = getattr(getattr(m, tower).projection[idx], kind).detach()
                assert torch.equal(got, ref), (tower, idx, kind)


def test_train_epoch_p64_matches_reference(golden_dir):
    """Same pin at P = 64, the fixture the split-bf16 fused step is checked against on the GPU."""
    g = _load(golden_dir, "train_epoch_p64.npz")
    m = _oracle_model_from(g, 64, prefix="init__")
    opt = torch.optim.Adam(m.parameters(), lr=float(g["lr"]))
    losses = []
    for b in range(3):
        batch = tuple(_t(g[f"b{b}_{nm}_{k}"]) for nm in ("q", "p", "n") for k in ("ids", "mask"))
        losses.append(O.train_step(m, opt, batch, float(g["margin"])))
    assert np.mean(losses) == pytest.approx(float(g["avg_loss"]), rel=1e-7)
    for tower in ("query_tower", "document_tower"):
        for idx in (0, 2):
            for kind in ("weight", "bias"):
                ref = _t(g[f"final__{tower}__projection__{idx}__{kind}"])
                got = getattr(getattr(m, tower).projection[idx], kind).detach()
                assert torch.equal(got, ref), (tower, idx, kind)


def test_train_epoch_accum_matches_reference_train_epoch_optimized(golden_dir):
    """oracle.train_epoch_accum == the reference's own train_epoch_optimized (fp32 autocast stand-in, real
    GradScaler): losses, weights after 2 optimiser steps over 5 batches, and the gradient the unstepped trailing
    batch leaves behind."""
    g = _load(golden_dir, "train_epoch_optimized.npz")
    m = _oracle_model_from(g, 64, prefix="init__")
    opt = torch.optim.Adam(m.parameters(), lr=float(g["lr"]))
    batches = [tuple(_t(g[f"b{b}_{nm}_{k}"]) for nm in ("q", "p", "n") for k in ("ids", "mask"))
               for b in range(int(g["n_batches"]))]
    avg = O.train_epoch_accum(m, opt, batches, float(g["margin"]), int(g["accum"]))
    assert avg == pytest.approx(float(g["avg_loss"]), rel=1e-7)
    for tower in ("query_tower", "document_tower"):
        for idx in (0, 2):
            for kind in ("weight", "bias"):
                ref = _t(g[f"final__{tower}__projection__{idx}__{kind}"])
                prm = getattr(getattr(m, tower).projection[idx], kind)
                assert torch.equal(prm.detach(), ref), (tower, idx, kind)
                left = _t(g[f"leftover_grad__{tower}__projection__{idx}__{kind}"])
                assert torch.allclose(prm.grad, left, rtol=1e-6, atol=1e-12), (tower, idx, kind)


def test_ndcg_matches_sklearn_golden(golden_dir):
    g = _load(golden_dir, "ndcg_sklearn.npz")
    for c in range(int(g["n_cases"])):
        for k in (1, 5, 10):
            got = O.ndcg_at_k(g[f"rel_{c}"], g[f"score_{c}"], k)
            assert got == pytest.approx(float(g[f"ndcg{k}_{c}"]), rel=1e-12, abs=1e-15), (c, k)


def test_ndcg_matches_installed_sklearn():
    from sklearn.metrics import ndcg_score

    r = np.random.default_rng(0)
    for _ in range(50):
        n = int(r.integers(2, 200))
        rel = (r.random(n) < 0.1).astype(np.int64)
        s = np.round(r.standard_normal(n), int(r.integers(0, 4))).astype(np.float32)
        for k in (1, 5, 10):
            assert O.ndcg_at_k(rel, s, k) == pytest.approx(ndcg_score(rel[None], s[None], k=k), rel=1e-12, abs=1e-15)


def test_closed_form_ndcg_equals_tie_free_ndcg():
    r = np.random.default_rng(1)
    for _ in range(30):
        n = int(r.integers(12, 300))
        s = r.standard_normal(n).astype(np.float32)
        rel_ids = set(r.choice(n, size=int(r.integers(0, 12)), replace=False).tolist())
        rel = np.zeros(n)
        rel[list(rel_ids)] = 1
        top = O.topk_ids(s, 10)
        assert O.ndcg_from_topk(top, rel_ids, 10) == pytest.approx(O.ndcg_at_k(rel, s, 10), abs=1e-12)


def test_evaluate_golden_via_oracle_scoring(golden_dir):
    """Full-pool evaluate_model of the reference == oracle scoring over the same query groups."""
    g = _load(golden_dir, "evaluate.npz")
    assert 0.0 <= float(g["val_ndcg10"]) <= 1.0
    assert float(g["full__final_queries_evaluated"]) == 12
