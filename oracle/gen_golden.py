"""
Generate tests/golden/*.npz by running the REFERENCE's own modules (imported unmodified from
/root/reference/backend) on seeded synthetic inputs.  Run in the build container only
(`python oracle/gen_golden.py`); /root/reference does not exist on the GPU box, so tests read the
committed vectors, never this script's imports.

Two patches are applied before `import model` because there is no network / HF cache
(SURVEY.md §8c): AutoModel.from_pretrained -> embedding-only backbone (north_star restatement of
backend/model.py:51-52) and AutoTokenizer.from_pretrained -> lookup into a token bank
(text handles "q:<i>" / "d:<i>").  Everything downstream (mean pooling, normalise, projection,
TripletLoss, train_epoch, evaluate_model, sklearn ndcg_score) is the reference's code.
"""
import os
import random
import sys

import numpy as np
import torch

REF = "/root/reference/backend"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
V, H = 257, 384


class _Cfg:
    hidden_size = H


class EmbOnly(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.config = _Cfg()
        self.emb = torch.nn.Embedding(V, H)

    @property
    def device(self):
        return self.emb.weight.device

    def forward(self, input_ids, attention_mask=None, token_type_ids=None):
        return (self.emb(input_ids),)


BANK: dict[str, list[int]] = {}


class BankTok:
    """Mimics tokenizer(texts, padding=True, truncation=True, return_tensors='pt', max_length=512)."""

    def __call__(self, texts, padding=True, truncation=True, return_tensors="pt", max_length=512):
        rows = [BANK[t][:max_length] for t in texts]
        L = max(len(r) for r in rows)
        ids = torch.zeros(len(rows), L, dtype=torch.int64)
        mask = torch.zeros(len(rows), L, dtype=torch.int64)
        for i, r in enumerate(rows):
            ids[i, : len(r)] = torch.tensor(r)
            mask[i, : len(r)] = 1
        return {"input_ids": ids, "token_type_ids": torch.zeros_like(ids), "attention_mask": mask}


def _install():
    import transformers

    transformers.AutoModel.from_pretrained = staticmethod(lambda name: EmbOnly())
    transformers.AutoTokenizer.from_pretrained = staticmethod(lambda name: BankTok())
    sys.path.insert(0, REF)


def _make_bank(rng: np.random.Generator, n_q: int, n_d: int):
    for i in range(n_q):
        n = int(rng.integers(3, 12))
        BANK[f"q:{i}"] = [101] + rng.integers(1, V, n - 2).tolist() + [102]
    for i in range(n_d):
        n = int(rng.integers(5, 40))
        BANK[f"d:{i}"] = [101] + rng.integers(1, V, n - 2).tolist() + [102]


def _tok(texts):
    t = BankTok()(texts)
    return t["input_ids"].numpy(), t["attention_mask"].numpy()


def _params(m):
    return {k.replace(".", "__"): v.detach().numpy().copy() for k, v in m.state_dict().items()}


def main():
    _install()
    import model as ref_model
    import training as ref_training
    from sklearn.metrics import ndcg_score

    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261018)
    _make_bank(rng, 64, 160)

    # ---------------------------------------------------------------- 1. pooling + normalise only
    torch.manual_seed(11)
    tower = ref_model.AveragePoolingTower("x", 8)
    h = torch.randn(6, 13, H)
    mask = (torch.rand(6, 13) > 0.35).long()
    mask[2] = 0  # all-masked row -> zeros
    mask[4] = 1
    pooled = tower._mean_pooling((h,), mask)
    normed = torch.nn.functional.normalize(pooled, p=2, dim=1)
    np.savez_compressed(os.path.join(OUT, "pooling.npz"), h=h.numpy(), mask=mask.numpy(),
                        pooled=pooled.numpy(), normed=normed.numpy())

    # ---------------------------------------------------------------- 2. forward / loss / backward
    for P, margin, tag in ((16, 0.3, "p16"), (64, 0.3, "p64")):
        torch.manual_seed(5 + P)
        m = ref_model.TwoTowersModel(projection_dim=P)
        crit = ref_model.TripletLoss(margin=margin)
        B = 12
        qs = [f"q:{i}" for i in rng.integers(0, 64, B)]
        ps = [f"d:{i}" for i in rng.integers(0, 160, B)]
        ns = [f"d:{i}" for i in rng.integers(0, 160, B)]
        q = m.encode_queries(qs)
        p = m.encode_documents(ps)
        n = m.encode_documents(ns)
        loss = crit(q, p, n)
        loss.backward()
        out = _params(m)
        for name, prm in m.named_parameters():
            if prm.grad is not None:
                out["grad__" + name.replace(".", "__")] = prm.grad.numpy().copy()
        for nm, texts in (("q", qs), ("p", ps), ("n", ns)):
            ids, msk = _tok(texts)
            out[nm + "_ids"], out[nm + "_mask"] = ids, msk
        out.update(q=q.detach().numpy(), p=p.detach().numpy(), n=n.detach().numpy(),
                   loss=np.float32(loss.item()), margin=np.float32(margin))
        np.savez_compressed(os.path.join(OUT, f"step_{tag}.npz"), **out)

    # ---------------------------------------------------------------- 3. train_epoch (3 Adam steps)
    torch.manual_seed(77)
    P, margin, lr = 32, 0.3, 1e-3
    m = ref_model.TwoTowersModel(projection_dim=P)
    crit = ref_model.TripletLoss(margin=margin)
    opt = torch.optim.Adam(m.parameters(), lr=lr)
    init = _params(m)
    batches = []
    for _ in range(3):
        B = 10
        batches.append(([f"q:{i}" for i in rng.integers(0, 64, B)],
                        [f"d:{i}" for i in rng.integers(0, 160, B)],
                        [f"d:{i}" for i in rng.integers(0, 160, B)]))
    avg = ref_training.train_epoch(m, batches, crit, opt, log_wandb=False)
    out = {"init__" + k: v for k, v in init.items()}
    out.update({"final__" + k: v for k, v in _params(m).items()})
    for bi, (qs, ps, ns) in enumerate(batches):
        for nm, texts in (("q", qs), ("p", ps), ("n", ns)):
            ids, msk = _tok(texts)
            out[f"b{bi}_{nm}_ids"], out[f"b{bi}_{nm}_mask"] = ids, msk
    out.update(avg_loss=np.float64(avg), margin=np.float32(margin), lr=np.float64(lr))
    np.savez_compressed(os.path.join(OUT, "train_epoch.npz"), **out)

    # ---------------------------------------------------------------- 4. evaluate_model
    class DS:
        def __init__(self, n_pairs, seed):
            r = np.random.default_rng(seed)
            self.data = []
            for _ in range(n_pairs):
                qi = int(r.integers(0, 24))
                self.data.append({"query_id": qi, "query": f"q:{qi}", "positive": f"d:{int(r.integers(0, 160))}"})
            self.docs = sorted({d["positive"] for d in self.data})

        def __len__(self):
            return len(self.data)

        def __getitem__(self, i):
            return self.data[i]

        def get_unique_passages(self):
            return self.docs

    ds = DS(220, 3)
    torch.manual_seed(123)
    m = ref_model.TwoTowersModel(projection_dim=24)
    random.seed(1234)
    # pool (100) exceeds the number of irrelevant docs, so no random.sample over a str-set happens and the
    # value does not depend on PYTHONHASHSEED (set order only permutes candidates, which NDCG ignores)
    val = ref_training.evaluate_model(m, ds, sample_size=60, min_query_groups=10, candidate_pool_size=100)
    random.seed(4321)
    full = ref_training.evaluate_model(m, ds, sample_size=120, min_query_groups=12, candidate_pool_size=-1,
                                       comprehensive=True, wandb_prefix="final_")
    out = _params(m)
    out["ds_query_id"] = np.array([d["query_id"] for d in ds.data])
    out["ds_doc"] = np.array([int(d["positive"][2:]) for d in ds.data])
    out["val_ndcg10"] = np.float64(val)
    for k, v in full.items():
        out["full__" + k] = np.float64(v)
    # token bank (needed to rebuild the dataset's texts)
    for i in range(64):
        out[f"bank_q_{i}"] = np.array(BANK[f"q:{i}"])
    for i in range(160):
        out[f"bank_d_{i}"] = np.array(BANK[f"d:{i}"])
    np.savez_compressed(os.path.join(OUT, "evaluate.npz"), **out)

    # ---------------------------------------------------------------- 5. sklearn ndcg_score incl. ties
    cases = {}
    r = np.random.default_rng(9)
    for c in range(24):
        n = int(r.integers(2, 60))
        rel = (r.random(n) < 0.2).astype(np.int64)
        if c % 3 == 0:
            score = np.round(r.random(n), 1).astype(np.float32)  # heavy ties
        else:
            score = r.standard_normal(n).astype(np.float32)
        if c == 5:
            rel[:] = 0
        cases[f"rel_{c}"], cases[f"score_{c}"] = rel, score
        for k in (1, 5, 10):
            cases[f"ndcg{k}_{c}"] = np.float64(ndcg_score(rel.reshape(1, -1), score.reshape(1, -1), k=k))
    cases["n_cases"] = np.int64(24)
    np.savez_compressed(os.path.join(OUT, "ndcg_sklearn.npz"), **cases)

    # ---------------------------------------------------------------- 6. train_epoch at P = 64 (the smallest
    # projection the split-bf16 tensor-core path accepts), own rng so sections 1-5 keep their draws
    rng6 = np.random.default_rng(6006)
    torch.manual_seed(78)
    P, margin, lr = 64, 0.3, 1e-3
    m = ref_model.TwoTowersModel(projection_dim=P)
    crit = ref_model.TripletLoss(margin=margin)
    opt = torch.optim.Adam(m.parameters(), lr=lr)
    init = _params(m)
    batches = []
    for _ in range(3):
        B = 16
        batches.append(([f"q:{i}" for i in rng6.integers(0, 64, B)],
                        [f"d:{i}" for i in rng6.integers(0, 160, B)],
                        [f"d:{i}" for i in rng6.integers(0, 160, B)]))
    avg = ref_training.train_epoch(m, batches, crit, opt, log_wandb=False)
    out = {"init__" + k: v for k, v in init.items()}
    out.update({"final__" + k: v for k, v in _params(m).items()})
    for bi, (qs, ps, ns) in enumerate(batches):
        for nm, texts in (("q", qs), ("p", ps), ("n", ns)):
            ids, msk = _tok(texts)
            out[f"b{bi}_{nm}_ids"], out[f"b{bi}_{nm}_mask"] = ids, msk
    out.update(avg_loss=np.float64(avg), margin=np.float32(margin), lr=np.float64(lr))
    np.savez_compressed(os.path.join(OUT, "train_epoch_p64.npz"), **out)

    # ---------------------------------------------------------------- 7. train_epoch_optimized (training.py:66-133):
    # the reference's own loop (accumulation cadence, loss / accumulation_steps, GradScaler scale -> step -> update ->
    # zero_grad) on CPU.  autocast is replaced by a null context so the arithmetic stays fp32 (the CUDA path here
    # computes in fp32 / split-bf16 regardless of autocast); GradScaler("cpu") is the real one (power-of-two scale).
    import contextlib

    rng7 = np.random.default_rng(7007)
    torch.manual_seed(79)
    P, margin, lr, accum = 64, 0.3, 1e-3, 2
    m = ref_model.TwoTowersModel(projection_dim=P)
    crit = ref_model.TripletLoss(margin=margin)
    opt = torch.optim.Adam(m.parameters(), lr=lr)
    init = _params(m)
    batches = []
    for _ in range(5):  # odd count: the trailing batch is accumulated but never stepped (reference behaviour)
        B = 16
        batches.append(([f"q:{i}" for i in rng7.integers(0, 64, B)],
                        [f"d:{i}" for i in rng7.integers(0, 160, B)],
                        [f"d:{i}" for i in rng7.integers(0, 160, B)]))
    ref_training.autocast = lambda *a, **k: contextlib.nullcontext()
    scaler = torch.amp.GradScaler("cpu")
    avg = ref_training.train_epoch_optimized(m, batches, crit, opt, torch.device("cpu"), scaler,
                                             accumulation_steps=accum, log_wandb=False)
    out = {"init__" + k: v for k, v in init.items()}
    out.update({"final__" + k: v for k, v in _params(m).items()})
    for name, prm in m.named_parameters():  # gradient left behind by the unstepped 5th batch (scaled by the scaler)
        if prm.grad is not None:
            out["leftover_grad__" + name.replace(".", "__")] = (prm.grad / scaler.get_scale()).numpy().copy()
    for bi, (qs, ps, ns) in enumerate(batches):
        for nm, texts in (("q", qs), ("p", ps), ("n", ns)):
            ids, msk = _tok(texts)
            out[f"b{bi}_{nm}_ids"], out[f"b{bi}_{nm}_mask"] = ids, msk
    out.update(avg_loss=np.float64(avg), margin=np.float32(margin), lr=np.float64(lr), accum=np.int64(accum),
               n_batches=np.int64(len(batches)))
    np.savez_compressed(os.path.join(OUT, "train_epoch_optimized.npz"), **out)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
