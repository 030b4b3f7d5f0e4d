"""CPU-side tests: the C-ABI library loads and exports every declared symbol (no compute), host logic of the
data feeder / tokenizer / sharding, and the N>1 plumbing over gloo with world_size 2."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as g

    g.build()
    import two_towers_overlords_b200 as pkg
    from two_towers_overlords_b200 import data, retrieval, training  # noqa: F401

    pkg.data, pkg.retrieval, pkg.training = data, retrieval, training
    return pkg


def test_library_exports_every_declared_symbol(pkg):
    header = open(os.path.join(ROOT, "include", "tt_b200.h")).read()
    declared = set(re.findall(r"\b(tt_[a-z0-9_]+)\s*\(", header))
    lib = pkg._native.load()
    assert declared == set(pkg._native.EXPORTED_SYMBOLS), declared ^ set(pkg._native.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.tt_version() == 100
    assert lib.tt_last_error() is not None


def test_library_has_sm100a_code_only(pkg):
    out = subprocess.run(["cuobjdump", "-lelf", pkg._native.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_gpu(pkg):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg._native.NativeError):
        pkg.ops.pool(torch.zeros(10, 384), torch.zeros(2, 3, dtype=torch.int64), None)
    with pytest.raises(RuntimeError):
        pkg.training.run_training(num_epochs=1, use_wandb=False)


def test_product_never_imports_the_oracle():
    impl = os.path.join(ROOT, "two-towers-overlords_b200")
    for dirpath, _, files in os.walk(impl):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "two_towers_oracle" not in text and "oracle/" not in text and "oracle." not in text, f


def test_hash_tokenizer_shape_and_determinism(pkg):
    from two_towers_overlords_b200.model import HashTokenizer

    tok = HashTokenizer()
    out = tok(["What is a two-tower model?", "short"], padding=True, truncation=True, return_tensors="pt", max_length=6)
    assert out["input_ids"].shape == out["attention_mask"].shape == (2, 6)
    assert out["input_ids"][0, 0] == 101 and out["input_ids"][1, 2] == 102 and out["attention_mask"][1, 3:].sum() == 0
    again = tok(["What is a two-tower model?"], max_length=6)
    assert torch.equal(again["input_ids"][0], out["input_ids"][0])
    assert int(out["input_ids"].max()) < 30522


def test_token_bank_batch_matches_padding_semantics(pkg):
    bank = pkg.data.TokenBank.synthetic("d", 50, "doc", seed=1)
    idx = np.array([3, 7, 11])
    tb = bank.batch(idx)
    lens = bank.lengths(idx)
    assert tb.input_ids.shape[1] == lens.max()
    for r, i in enumerate(idx):
        assert np.array_equal(tb.input_ids[r, : lens[r]].numpy(), bank.tokens(int(i)))
        assert tb.attention_mask[r].sum() == lens[r] and tb.input_ids[r, lens[r]:].sum() == 0
        assert tb.input_ids[r, 0] == 101 and tb.input_ids[r, lens[r] - 1] == 102
    fixed = bank.batch(idx, max_length=32, pad_to=32, ids_dtype=torch.int32, mask_dtype=torch.uint8)
    assert fixed.input_ids.shape == (3, 32) and fixed.input_ids.dtype == torch.int32
    tok = pkg.data.TokenBankTokenizer(bank)
    out = tok([bank.handle(3), bank.handle(7)])
    assert torch.equal(out["input_ids"], bank.batch(np.array([3, 7])).input_ids)


def test_synthetic_dataset_shape(pkg):
    ds = pkg.data.MSMarcoDataset("train", max_samples=500, synthetic=True)
    assert len(ds) == 500 and len(ds.get_unique_passages()) == 500
    item = ds[0]
    assert set(item) == {"query_id", "query", "positive"}
    per_q = np.bincount([d["query_id"] for d in ds.data])
    assert 4 < per_q.mean() < 12  # ~8.2 passages per query like MS MARCO
    with pytest.raises(ValueError):
        pkg.data.MSMarcoDataset("train", max_samples=0, synthetic=True)


def test_in_batch_negatives_follow_reference_rule(pkg):
    rng = np.random.default_rng(0)
    qids = rng.integers(0, 40, 256)
    for _ in range(20):
        j = pkg.data.sample_negative_indices(qids, rng)
        assert (j != np.arange(256)).all() and (qids[j] != qids).all()
    with pytest.raises(ValueError):
        pkg.data.sample_negative_indices(np.zeros(8, dtype=np.int64), rng)
    ds = pkg.data.MSMarcoDataset("train", max_samples=300, synthetic=True)
    dl = pkg.data.TripletDataLoader(ds, batch_size=64, seed=1)
    seen = 0
    pos_to_q = {d["positive"]: d["query_id"] for d in ds.data}
    q_to_id = {d["query"]: d["query_id"] for d in ds.data}
    for queries, positives, negatives in dl:
        assert len(queries) == len(positives) == len(negatives)
        for q, p, n in zip(queries, positives, negatives):
            assert n != p and pos_to_q[n] != q_to_id[q]
        seen += len(queries)
    assert seen == 300 and len(dl) == 5


def test_token_triplet_loader_shards_the_global_batch(pkg):
    ds = pkg.data.MSMarcoDataset("train", max_samples=512, synthetic=True)
    full = list(pkg.data.TokenTripletLoader(ds, 128, 32, 256, rank=0, world_size=1, seed=5))
    r0 = list(pkg.data.TokenTripletLoader(ds, 128, 32, 256, rank=0, world_size=2, seed=5))
    r1 = list(pkg.data.TokenTripletLoader(ds, 128, 32, 256, rank=1, world_size=2, seed=5))
    assert len(full) == len(r0) == len(r1) == 4
    for f, a, b in zip(full, r0, r1):
        for k in range(3):
            assert torch.equal(f[k].input_ids, torch.cat([a[k].input_ids, b[k].input_ids]))
            assert torch.equal(f[k].attention_mask, torch.cat([a[k].attention_mask, b[k].attention_mask]))
        assert f[0].input_ids.shape == (128, 32) and f[1].input_ids.shape == (128, 256)
        assert f[0].input_ids.dtype == torch.int32 and f[0].attention_mask.dtype == torch.uint8


def test_shard_bounds_cover_the_corpus(pkg):
    for n in (0, 1, 7, 1000, 8_800_001):
        for w in (1, 2, 4, 8):
            spans = [pkg.retrieval.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_relevance_csr(pkg):
    offs, flat = pkg.retrieval.relevance_csr([{5, 2}, set(), {9}], "cpu")
    assert offs.tolist() == [0, 2, 2, 3] and flat.tolist() == [2, 5, 9]


def test_tie_averaged_ndcg_matches_sklearn(pkg):
    from sklearn.metrics import ndcg_score

    r = np.random.default_rng(0)
    for _ in range(40):
        n = int(r.integers(2, 80))
        rel = (r.random(n) < 0.2).astype(np.int64)
        s = np.round(r.standard_normal(n), int(r.integers(0, 3))).astype(np.float32)
        for k in (1, 5, 10):
            assert pkg.training._tie_averaged_ndcg(rel, s, k) == pytest.approx(ndcg_score(rel[None], s[None], k=k), abs=1e-12)


def test_ndcg_from_lists_handles_ties_inside_the_list(pkg):
    from sklearn.metrics import ndcg_score

    r = np.random.default_rng(3)
    for _ in range(40):
        n = 60
        rel = (r.random(n) < 0.25).astype(np.int64)
        s = np.round(r.standard_normal(n), 1).astype(np.float32)
        order = np.lexsort((np.arange(n), -s.astype(np.float64)))[:16]
        fallback = lambda: pkg.training._tie_averaged_ndcg(rel, s, 10)  # noqa: E731
        got = pkg.training._ndcg_from_lists(s[order], order.astype(np.int64), set(np.flatnonzero(rel).tolist()),
                                            int(rel.sum()), 10, fallback)
        assert got == pytest.approx(ndcg_score(rel[None], s[None], k=10), abs=1e-12)


_GLOO_WORKER = r"""
import os, sys, torch, numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["TT_ROOT"])
from two_towers_overlords_b200 import data, retrieval, training
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["TT_PORT"], rank=int(os.environ["RANK"]), world_size=2)
rank = dist.get_rank()
# 1. every rank draws the same global batches and owns a disjoint contiguous slice
ds = data.MSMarcoDataset("train", max_samples=256, synthetic=True)
q, p, n = next(iter(data.TokenTripletLoader(ds, 64, 32, 256, rank=rank, world_size=2, seed=3)))
mine = q.input_ids.to(torch.int64)
both = [torch.empty_like(mine) for _ in range(2)]
dist.all_gather(both, mine)
full = next(iter(data.TokenTripletLoader(ds, 64, 32, 256, rank=0, world_size=1, seed=3)))[0].input_ids.to(torch.int64)
assert torch.equal(torch.cat(both), full)
# 2. gradient all-reduce of the flat buffer (+ loss slot): sum of per-rank partial means == global mean
flat = torch.full((11,), float(rank + 1)); flat[-1] = 0.25 * (rank + 1)
dist.all_reduce(flat)
assert torch.allclose(flat[:-1], torch.full((10,), 3.0)) and abs(float(flat[-1]) - 0.75) < 1e-6
# 3. corpus shards: all-gather of per-shard top-k lists reassembles [G,Q,k]
lo, hi = retrieval.shard_bounds(1001, 2, rank)
s = torch.full((4, 10), float(rank)); i = torch.arange(lo, lo + 40).view(4, 10)
ps, pi = retrieval.all_gather_lists(s, i, 2)
assert ps.shape == (2, 4, 10) and pi.dtype == torch.int64
assert ps[1].eq(1).all() and int(pi[1, 0, 0]) == retrieval.shard_bounds(1001, 2, 1)[0]
# 4. evaluate_model under data parallelism: rank 0's sampling plan becomes everybody's (the ranks' `random` states
#    differ), and every rank owns a contiguous shard of the document universe
import random
random.seed(100 + rank)
val = data.MSMarcoDataset("validation", max_samples=300, synthetic=True)
box = [training._eval_plan(val, 60, 8, 20, say=lambda *a, **k: None) if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
plan = box[0]
sig = repr((plan["queries"], plan["rel_docs"], plan["cand_docs"]))
sigs = [None, None]
dist.all_gather_object(sigs, sig)
assert sigs[0] == sigs[1] and len(plan["queries"]) == 8 and all(len(c) >= 2 for c in plan["cand_docs"])
universe = list(dict.fromkeys(d for r in plan["rel_docs"] for d in r))
lo, hi = retrieval.shard_bounds(len(universe), 2, rank)
spans = [None, None]
dist.all_gather_object(spans, (lo, hi))
assert spans[0][0] == 0 and spans[0][1] == spans[1][0] and spans[1][1] == len(universe)
# 5. shard embeddings of different heights reassemble in rank order (validation pools gather the shards)
rows = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 3)
allrows = training._gather_rows(rows, 2)
assert allrows.shape == (len(universe), 3) and torch.equal(allrows[:, 0], torch.arange(len(universe), dtype=torch.float32))
# 6. mixed-bank handles resolve per bank (train + validation passages in one call)
tok = val.tokenizer()
tok.add(ds.tokenizer())
mixed = tok([val.docs[0], ds.docs[0], val.docs[1]])
assert mixed["input_ids"].shape[0] == 3
assert torch.equal(mixed["input_ids"][1, : int(mixed["attention_mask"][1].sum())].to(torch.int64),
                   torch.from_numpy(ds.doc_bank.tokens(0)).to(torch.int64))
dist.destroy_process_group()
print("ok", rank)
"""


def test_world_size_2_plumbing_over_gloo(pkg, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), TT_ROOT=ROOT, TT_PORT=port, HF_HUB_OFFLINE="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_search_model_discovery_follows_reference_rules(pkg, tmp_path):
    """find_best_model / get_projection_dim_from_model (backend/search.py:41-117): most epochs wins, weights.pt
    overrides, projection dim comes from query_tower.projection.2.weight."""
    import torch

    from two_towers_overlords_b200 import search

    assert search.find_best_model(str(tmp_path / "missing")) is None
    assert search.find_best_model(str(tmp_path)) is None
    for name, P in (("e3.lr3.d64.m3.pt", 64), ("e15.lr4.d384.m3.pt", 384), ("notes.txt", 0)):
        if P:
            torch.save({"query_tower.projection.2.weight": torch.zeros(P, P)}, tmp_path / name)
        else:
            (tmp_path / name).write_text("x")
    path, fname = search.find_best_model(str(tmp_path))
    assert fname == "e15.lr4.d384.m3.pt" and search.get_projection_dim_from_model(path) == 384
    torch.save({"other": torch.zeros(1)}, tmp_path / "weights.pt")
    path, fname = search.find_best_model(str(tmp_path))
    assert fname == "weights.pt" and search.get_projection_dim_from_model(path) == search.DEFAULT_PROJ_DIM


def test_ctypes_signatures_match_the_header_prototypes(pkg):
    """Every prototype in include/tt_b200.h has a ctypes signature with the same number of parameters (binding
    drift shows up here, not as a crash on the GPU box)."""
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "tt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = dict(re.findall(r"\b(tt_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text))
    sigs = pkg._native._SIGNATURES
    assert set(protos) == set(sigs), set(protos) ^ set(sigs)
    for name, params in protos.items():
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(sigs[name][1]), (name, n, len(sigs[name][1]))


def test_step_args_struct_matches_the_header(pkg):
    """tt_step_args: the ctypes mirror lists the same fields in the same order as the C struct."""
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "tt_b200.h")).read()
    body = re.search(r"typedef struct tt_step_args \{(.*?)\} tt_step_args;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", " ", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            names.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", part)[-1])
    assert names == [f[0] for f in pkg._native.StepArgs._fields_]


def test_reference_arm_prints_one_json_line_on_cpu():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) works without a GPU."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_triplets_per_sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_bench_synthetic_token_sets_follow_the_stated_shape():
    """bench.py's synthetic inputs: shape U (full-length rows, ids in [999, 30522)), negatives = the positives of
    another item of the batch (no fixed point), same bytes for the same seed."""
    import importlib.util

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("tt_bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    a = bench.token_sets(2, 64, 7, torch.uint16, torch.uint8, pin=False)
    b = bench.token_sets(2, 64, 7, torch.uint16, torch.uint8, pin=False)
    for sa, sb in zip(a, b):
        q, qm, p, pm, n, nm = sa
        assert tuple(q.shape) == (64, bench.LQ) and tuple(p.shape) == (64, bench.LD) and q.dtype == torch.uint16
        assert int(q.to(torch.int64).min()) >= 999 and int(p.to(torch.int64).max()) < bench.VOCAB
        assert bool(qm.all()) and bool(pm.all()) and bool(nm.all())
        pi, ni = p.to(torch.int64), n.to(torch.int64)
        assert not any(torch.equal(pi[i], ni[i]) for i in range(64))          # never its own positive
        assert all(any(torch.equal(ni[i], pi[j]) for j in range(64)) for i in range(0, 64, 9))  # an in-batch positive
        assert all(torch.equal(x, y) for x, y in zip(sa, sb))
    cfg = bench.workload_config(8)
    assert cfg["global_batch"] == 8 * bench.B_PER_GPU and cfg["parallelism"] == "dp8" and "workload" in cfg
