"""GPU parity of the WHOLE training step on the code path bench.py times (FusedTrainer -> tt_triplet_step, split-bf16
tensor-core projection) against the CPU oracle at BASELINE.json's sizes: configs[1] (B = 2048, P = 512, 32 / 256
tokens) and configs[2] (B = 4096, P = 384, trainable 30522 x 384 tables).  Gates are BASELINE.json's: pooled rows and
loss 1e-4 relative, every gradient 1e-3 relative.  Also the reference's accumulation loop (training.py:66-133)."""
import os

import numpy as np
import pytest
import torch

from oracle import two_towers_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda"
PRECISIONS = [p for p in os.environ.get("TT_TEST_PRECISIONS", "fp32,bf16x3").split(",") if p]
NAMES = ("query_tower.projection.0.weight", "query_tower.projection.0.bias", "query_tower.projection.2.weight",
         "query_tower.projection.2.bias", "document_tower.projection.0.weight", "document_tower.projection.0.bias",
         "document_tower.projection.2.weight", "document_tower.projection.2.bias")


@pytest.fixture(scope="module")
def tt():
    import two_towers_overlords_b200 as pkg
    from two_towers_overlords_b200 import data, ops, training

    pkg.ops, pkg.training, pkg.data = ops, training, data
    return pkg


@pytest.fixture(params=["kernels", "persistent"])
def chain(request, monkeypatch):
    """Both ways of launching the projection chain (tt_step_args.chain): one kernel per contraction, and the persistent
    chain kernel (loss in the layer-2 epilogue, optimiser in its tail).  fp32 ignores the switch."""
    monkeypatch.setenv("TT_CHAIN", "1" if request.param == "persistent" else "0")
    return request.param


def rel_err(got, want) -> float:
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return float((got - want).norm() / want.norm().clamp_min(1e-30))


def row_rel(got, want, floor=1e-6) -> float:
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return float(((got - want).norm(dim=1) / want.norm(dim=1).clamp_min(floor)).max())


def oracle_twin(m, P, vocab, train_table=False):
    ref = O.OracleTwoTowers(P, vocab=vocab, train_table=train_table)
    with torch.no_grad():
        for t_new, t_ref in ((m.query_tower, ref.query_tower), (m.document_tower, ref.document_tower)):
            t_ref.table.copy_(t_new.pretrained_model.table.detach().float().cpu())
            for i in (0, 2):
                t_ref.projection[i].weight.copy_(t_new.projection[i].weight.cpu())
                t_ref.projection[i].bias.copy_(t_new.projection[i].bias.cpu())
    return ref


KINK_BAND = 2e-6  # |pre-activation| below which the ReLU gate is undecided between two fp32 summation orders


def oracle_grads(ref, batch, margin, gate=None):
    """Loss and the 8 projection gradients of the oracle (+ table gradients when they train).  Also returns the
    first-layer pre-activations z [3B,P] (rows q | p | n).

    `gate` (bool [3B,P], optional): the ReLU gate to differentiate through instead of the oracle's own `z > 0`.
    ReLU is not differentiable at 0: an element whose pre-activation lies within rounding error of zero (|z| of a few
    1e-8 against a scale of 3e-2) gets gate 1 from one correct fp32 summation order and gate 0 from another — the
    CPU oracle itself flips ~0.4 such elements per configs[1] step against float64.  ONE flipped element moves a
    weight gradient by ~1e-3 of its norm (its rank-1 term is as large as any other of the 3.1 M, and the gradient is
    a sum of random signs), so the 1e-3 gradient gate is only meaningful with the same gate on both sides — exactly
    like the top-10 id rule, which is exact "wherever score gaps exceed 1e-5".  The tests therefore (1) require every
    gate disagreement to sit inside |z| <= KINK_BAND, (2) count them, and (3) compare gradients with the oracle
    differentiated through the device's gate; the forward value is unchanged by more than KINK_BAND."""
    for prm in ref.parameters():
        prm.grad = None
    ys, zs = [], []
    B = batch.q_ids.shape[0]
    for g, (tower, ids, mask) in enumerate(((ref.query_tower, batch.q_ids, batch.q_mask),
                                            (ref.document_tower, batch.p_ids, batch.p_mask),
                                            (ref.document_tower, batch.n_ids, batch.n_mask))):
        if tower.table.requires_grad:
            emb = torch.nn.functional.embedding(ids.long(), tower.table)
        else:
            with torch.no_grad():
                emb = torch.nn.functional.embedding(ids.long(), tower.table)
        x = torch.nn.functional.normalize(O.mean_pooling(emb, mask.long()), p=2, dim=1)
        z = torch.nn.functional.linear(x, tower.projection[0].weight, tower.projection[0].bias)
        h = torch.relu(z) if gate is None else z * gate[g * B: (g + 1) * B].to(z.dtype)
        ys.append(torch.nn.functional.linear(h, tower.projection[2].weight, tower.projection[2].bias))
        zs.append(z.detach())
    loss = O.triplet_loss(ys[0], ys[1], ys[2], margin)
    loss.backward()
    grads = [dict(ref.named_parameters())[nm].grad.clone() for nm in NAMES]
    return float(loss.item()), grads, torch.cat(zs)


def check_gate_and_grads(tr, ref, batch, margin, tag):
    """Shared gradient check of the fused step (see oracle_grads): returns the number of ReLU kink flips."""
    ref_loss, raw_grads, z = oracle_grads(ref, batch, margin)
    gate = tr.step_obj.relu_gate().cpu()
    flips = gate != (z > 0)
    n_flips = int(flips.sum())
    assert n_flips <= 16, (tag, n_flips)                      # a handful of 3.1 M, not a systematic gate error
    assert float(z[flips].abs().max()) <= KINK_BAND if n_flips else True, (tag, z[flips])
    _, grads, _ = oracle_grads(ref, batch, margin, gate=gate)  # same gate on both sides
    for nm, got, want_g, raw in zip(NAMES, tr.g_views, grads, raw_grads):
        assert rel_err(got, want_g) < 1e-3, (tag, nm, rel_err(got, want_g))
        assert rel_err(got, raw) < 5e-3 * max(1, n_flips), (tag, nm, rel_err(got, raw), n_flips)  # never far off the raw one
    return ref_loss, n_flips


@pytest.mark.parametrize("shape", ["U", "Z"])
@pytest.mark.parametrize("precision", PRECISIONS)
def test_fused_step_configs1_vs_oracle(tt, precision, shape, chain):
    """configs[1] in full: one fused step (the C path bench.py's headline runs) against oracle.train_step's
    forward/backward: loss, all 6144 pooled rows, all 8 projection gradients."""
    if precision == "fp32" and chain == "persistent":
        pytest.skip("the fp32 CUDA-core path has no chain kernel")
    torch.manual_seed(0)
    B, Lq, Ld, P, margin = 2048, 32, 256, 512, 0.3
    m = tt.TwoTowersModel(projection_dim=P, precision=precision).to(DEV)
    batch = O.synth_triplet_batch(B, Lq, Ld, shape, seed=5 if shape == "U" else 6)
    Lq_, Ld_ = batch.q_ids.shape[1], batch.p_ids.shape[1]
    tr = tt.training.FusedTrainer(m, margin, 1e-3, B, Lq_, Ld_, precision=precision, use_graph=False,
                                  ids_dtype=torch.int32, mask_dtype=torch.uint8)
    for dst, src in zip(tr.tok, batch.astuple()):
        dst.copy_(src.to(dst.dtype))
    tr._fwd_bwd()
    torch.cuda.synchronize()
    ref = oracle_twin(m, P, O.VOCAB)
    ref_loss, n_flips = check_gate_and_grads(tr, ref, batch, margin, (precision, shape))
    got_loss = float(tr.loss_view.item())
    assert abs(got_loss - ref_loss) <= 1e-4 * abs(ref_loss), (got_loss, ref_loss)
    # pooled + normalised rows, q | p | n
    xhat = tr.step_obj.pooled_rows().cpu()
    want = torch.cat([O.pooled_normalised(ref.query_tower.table, batch.q_ids, batch.q_mask),
                      O.pooled_normalised(ref.document_tower.table, batch.p_ids, batch.p_mask),
                      O.pooled_normalised(ref.document_tower.table, batch.n_ids, batch.n_mask)])
    assert row_rel(xhat, want) < 1e-4
    print(f"configs[1] {precision} {shape}: {n_flips} ReLU kink flips of {3 * B * P}")


@pytest.mark.parametrize("precision", PRECISIONS)
def test_fused_step_configs2_trainable_table_vs_oracle(tt, precision, chain):
    """configs[2] in full (saved-model shape): B = 4096, P = 384, both 30522 x 384 tables trainable — projection
    gradients and both table gradients (sorted-segment scatter-add) against autograd through nn.Embedding."""
    if precision == "fp32" and chain == "persistent":
        pytest.skip("the fp32 CUDA-core path has no chain kernel")
    torch.manual_seed(2)
    B, Lq, Ld, P, margin = 4096, 32, 256, 384, 0.3
    m = tt.TwoTowersModel(projection_dim=P, precision=precision, train_table=True).to(DEV)
    batch = O.synth_triplet_batch(B, Lq, Ld, "Z", seed=12)
    Lq_, Ld_ = batch.q_ids.shape[1], batch.p_ids.shape[1]
    tr = tt.training.FusedTrainer(m, margin, 1e-3, B, Lq_, Ld_, precision=precision, use_graph=False,
                                  ids_dtype=torch.int32, mask_dtype=torch.uint8)
    for dst, src in zip(tr.tok, batch.astuple()):
        dst.copy_(src.to(dst.dtype))
    tr._fwd_bwd()
    torch.cuda.synchronize()
    first = [g.clone() for g in tr.table_grads]
    ref = oracle_twin(m, P, O.VOCAB, train_table=True)
    ref_loss, n_flips = check_gate_and_grads(tr, ref, batch, margin, (precision, "configs2"))
    assert abs(float(tr.loss_view.item()) - ref_loss) <= 1e-4 * abs(ref_loss)
    # Table gradients: rows named by tens of thousands of tokens (Zipf heads: 74 k terms in one row here) are long
    # sums of random signs, where the fp32 CPU oracle ITSELF is 1.3e-3 off float64 (sequential index_add) while the
    # device's fixed-order split sums are 3e-4 off (scripts/diag_table_grad.py) — so the truth for this D2 extension
    # is float64 autograd through nn.Embedding, differentiated through the device's ReLU gate like the rest.
    ref64 = oracle_twin(m, P, O.VOCAB, train_table=True).double()
    oracle_grads(ref64, batch, margin, gate=tr.step_obj.relu_gate().cpu())
    for got, tower in zip(tr.table_grads, (ref64.query_tower, ref64.document_tower)):
        assert rel_err(got, tower.table.grad) < 1e-3
        untouched = tower.table.grad.abs().sum(1) == 0
        assert float(got.cpu()[untouched].abs().max()) == 0.0  # rows no token named stay exactly zero
    # deterministic: the scatter-add has a fixed summation order
    tr._fwd_bwd()
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(first, tr.table_grads))


def _golden_model(tt, g, P, precision):
    V = g["init__query_tower__pretrained_model__emb__weight"].shape[0]
    m = tt.TwoTowersModel(projection_dim=P, vocab_size=V, precision=precision)
    with torch.no_grad():
        for name in ("query_tower", "document_tower"):
            t = getattr(m, name)
            t.pretrained_model.table.copy_(torch.from_numpy(g[f"init__{name}__pretrained_model__emb__weight"]))
            for i in (0, 2):
                t.projection[i].weight.copy_(torch.from_numpy(g[f"init__{name}__projection__{i}__weight"]))
                t.projection[i].bias.copy_(torch.from_numpy(g[f"init__{name}__projection__{i}__bias"]))
    return m.to(DEV)


def _golden_batches(g, n):
    return [tuple((torch.from_numpy(g[f"b{b}_{nm}_ids"]), torch.from_numpy(g[f"b{b}_{nm}_mask"]))
                  for nm in ("q", "p", "n")) for b in range(n)]


def _check_updates(g, m, tol, tag):
    for name in ("query_tower", "document_tower"):
        for i in (0, 2):
            for kind in ("weight", "bias"):
                want = torch.from_numpy(g[f"final__{name}__projection__{i}__{kind}"])
                init = torch.from_numpy(g[f"init__{name}__projection__{i}__{kind}"])
                got = getattr(getattr(m, name).projection[i], kind).detach().cpu()
                # the UPDATE (a few Adam steps of size ~lr), not the weights, so the check has teeth
                assert rel_err(got - init, want - init) < tol, (tag, name, i, kind, rel_err(got - init, want - init))


@pytest.mark.parametrize("mode", ["train_epoch", "fused", "fused_graph"])
@pytest.mark.parametrize("precision", PRECISIONS)
def test_three_adam_steps_p64_match_reference_train_epoch(tt, golden_dir, mode, precision, chain):
    """The reference's own train_epoch (3 Adam steps, P = 64) against every way of running the step here, in both
    arithmetic modes — the split-bf16 fused step is the one bench.py times."""
    if (precision == "fp32" or mode == "train_epoch") and chain == "persistent":
        pytest.skip("only the fused tensor-core step has a chain kernel")
    g = np.load(os.path.join(golden_dir, "train_epoch_p64.npz"))
    m = _golden_model(tt, g, 64, precision)
    margin, lr = float(g["margin"]), float(g["lr"])
    batches = _golden_batches(g, 3)
    if mode == "train_epoch":
        opt = torch.optim.Adam(m.parameters(), lr=lr)
        avg = tt.training.train_epoch(m, batches, tt.TripletLoss(margin), opt, log_wandb=False)
    else:
        B = batches[0][0][0].shape[0]
        Lq = max(b[0][0].shape[1] for b in batches)
        Ld = max(max(b[1][0].shape[1], b[2][0].shape[1]) for b in batches)
        pad = lambda t, L: torch.nn.functional.pad(t, (0, L - t.shape[1]))  # noqa: E731
        tr = tt.training.FusedTrainer(m, margin, lr, B, Lq, Ld, precision=precision, use_graph=(mode == "fused_graph"),
                                      ids_dtype=torch.int64, mask_dtype=torch.int64)
        losses = []
        for (q, p, n) in batches:
            srcs = (pad(q[0], Lq), pad(q[1], Lq), pad(p[0], Ld), pad(p[1], Ld), pad(n[0], Ld), pad(n[1], Ld))
            for dst, src in zip(tr.tok, srcs):
                dst.copy_(src)
            losses.append(float(tr.step().item()))
        avg = float(np.mean(losses))
    assert abs(avg - float(g["avg_loss"])) <= 1e-4 * float(g["avg_loss"])
    _check_updates(g, m, 2e-3, (mode, precision))


@pytest.mark.parametrize("precision", PRECISIONS)
def test_train_epoch_optimized_matches_reference_golden(tt, golden_dir, precision):
    """train_epoch_optimized (training.py:66-133): autocast + GradScaler + accumulation of 2 over 5 batches, against
    the reference's own function (golden): mean loss, the weights after the 2 optimiser steps, and the (unscaled)
    gradient the trailing unstepped batch leaves in .grad."""
    g = np.load(os.path.join(golden_dir, "train_epoch_optimized.npz"))
    m = _golden_model(tt, g, 64, precision)
    margin, lr, accum = float(g["margin"]), float(g["lr"]), int(g["accum"])
    batches = _golden_batches(g, int(g["n_batches"]))
    opt = torch.optim.Adam(m.parameters(), lr=lr)
    scaler = torch.amp.GradScaler("cuda")
    avg = tt.training.train_epoch_optimized(m, batches, tt.TripletLoss(margin), opt, torch.device("cuda"), scaler,
                                            accumulation_steps=accum, log_wandb=False)
    assert abs(avg - float(g["avg_loss"])) <= 1e-4 * float(g["avg_loss"])
    _check_updates(g, m, 2e-3, ("optimized", precision))
    scale = scaler.get_scale()
    for name in ("query_tower", "document_tower"):
        for i in (0, 2):
            for kind in ("weight", "bias"):
                want = torch.from_numpy(g[f"leftover_grad__{name}__projection__{i}__{kind}"])
                got = getattr(getattr(m, name).projection[i], kind).grad / scale
                assert rel_err(got, want) < 1e-3, (name, i, kind)
    # cadence: 5 batches, accumulation 2 -> exactly 2 optimiser steps
    assert int(next(iter(opt.state.values()))["step"]) == 2
