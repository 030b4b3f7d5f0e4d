"""
Document index + search — B200-native drop-in for the reference's backend/search.py surface
(`find_best_model`, `get_projection_dim_from_model`, `DocumentSearchEngine`, `build_document_index`,
`search_documents`; backend/search.py:41-117,154-542, consumed by backend/api.py:131-166).

The reference encodes documents with the document tower, ships fp32 `embedding.tobytes()` records to a Redis
HNSW index (approximate, cosine) and answers a query with one VectorQuery (search.py:268-402).  Here the index is
the corpus store of the scan kernel: normalised fp32 rows (+ bf16 copies) resident in HBM, answered EXACTLY by
tt_scan_topk (tensor-core candidates + exact fp32 re-score), so no Redis is needed; the Redis arguments are
accepted and ignored so callers keep working.  Persistence (Redis kept the index across processes) is a
`.npz` beside the model files: the same fp32 embedding bytes + the document texts.

Result format = the reference's (search.py:385-401): `{"id", "content", "score", "distance"}` with
`score = (1 + cos) / 2` (RedisVL `normalize_vector_distance=True` maps the cosine distance 1 - cos from [0, 2]
to a similarity in [0, 1]) and `distance = 1 - score`.
"""
from __future__ import annotations

import os
import random
import re
from typing import Any, Optional

import numpy as np
import torch

try:
    from . import ops, retrieval
    from .data import MSMarcoDataset
    from .model import TwoTowersModel
except ImportError:
    import ops
    import retrieval
    from data import MSMarcoDataset
    from model import TwoTowersModel

MODELS_DIR = "../models"
WEIGHTS_OVERRIDE = "weights.pt"
DEFAULT_INDEX_NAME = "default_index"
DEFAULT_PROJ_DIM = 128


def find_best_model(models_dir: str = MODELS_DIR) -> Optional[tuple[str, str]]:
    """Model file with the most epochs (`e{E}.lr{k}.d{P}.m{m}.pt`, main.py:12-13), unless `weights.pt` is present
    (search.py:41-84).  Returns (path, filename) or None."""
    if not os.path.isdir(models_dir):
        print(f"❌ Models directory '{models_dir}' does not exist or is not a directory")
        return None
    found = []
    for filename in os.listdir(models_dir):
        if filename.endswith(".pt"):
            m = re.match(r"^e(\d+)", filename)
            if m:
                found.append((int(m.group(1)), os.path.join(models_dir, filename), filename))
            elif filename == WEIGHTS_OVERRIDE:
                print(f"🔍 Best model selection overridden by presence of {WEIGHTS_OVERRIDE} file in {models_dir}")
                return os.path.join(models_dir, filename), filename
    if not found:
        print(f"ℹ️ No trained models found in '{models_dir}'")
        return None
    found.sort(key=lambda x: x[0], reverse=True)
    epochs, path, filename = found[0]
    print(f"🎯 Auto-selected best model: {filename} (trained for {epochs} epochs)")
    return path, filename


def get_projection_dim_from_model(model_path: str) -> int:
    """Projection dim = rows of `query_tower.projection.2.weight` in the saved state dict (search.py:87-117)."""
    try:
        sd = torch.load(model_path, map_location="cpu", weights_only=True)
        if "query_tower.projection.2.weight" in sd:
            dim = int(sd["query_tower.projection.2.weight"].shape[0])
            print(f"🔍 Detected projection dimension: {dim}")
            return dim
        print("⚠️ Could not find projection layer in model state dict, using default dimension")
        return DEFAULT_PROJ_DIM
    except Exception as e:  # noqa: BLE001  (the reference swallows load errors the same way)
        print(f"⚠️ Error reading model file {model_path}: {e}")
        return DEFAULT_PROJ_DIM


class DocumentSearchEngine:
    """Exact vector search over documents held in this GPU's HBM (replaces the Redis HNSW index)."""

    def __init__(self, model_filename: Optional[str] = None, models_dir: str = MODELS_DIR,
                 projection_dim: Optional[int] = None, redis_url: str = "redis://localhost:6379",
                 redis_host: Optional[str] = None, redis_port: Optional[int] = 6379, redis_pass: Optional[str] = None,
                 index_dir: Optional[str] = None, precision: str = "bf16", model: Optional[TwoTowersModel] = None,
                 **model_kwargs):
        model_path = actual = None
        if model is None:
            if model_filename is None:
                res = find_best_model(models_dir)
                if res:
                    model_path, actual = res
            elif model_filename.endswith((".pt", ".pth")):
                path = os.path.join(models_dir, model_filename)
                if os.path.exists(path):
                    model_path, actual = path, model_filename
                print(f"🔍 Using custom model: {model_filename}")
            else:
                print("❌ Custom model provided but either not found to exist, or not ending with .pt/.pth")
        self.index_name = os.path.splitext(actual)[0] if actual else DEFAULT_INDEX_NAME
        if model is not None:
            self.projection_dim = model.query_tower.projection[2].out_features
        elif projection_dim is None:
            self.projection_dim = get_projection_dim_from_model(model_path) if model_path else DEFAULT_PROJ_DIM
        else:
            self.projection_dim = projection_dim
        self.model = model if model is not None else TwoTowersModel(projection_dim=self.projection_dim, **model_kwargs)
        if model_path:
            print(f"🔍 Loading model weights from: {model_path}")
            self.model.load_state_dict(torch.load(model_path, map_location="cpu", weights_only=True), strict=False)
        elif model is None:
            print("ℹ️ No model weights provided or found - using untrained model with random weights")
        self.model.eval()
        if not torch.cuda.is_available():
            raise RuntimeError("DocumentSearchEngine needs a CUDA (sm_100a) device: the scan has no CPU fallback")
        self.device = torch.device("cuda")
        self.model = self.model.to(self.device)
        self.precision = precision
        self.index_dir = index_dir if index_dir is not None else models_dir
        self.documents: list[str] = []
        self.embeddings: Optional[torch.Tensor] = None  # [N, P] fp32 (device): what Redis stored as tobytes()
        self.shard: Optional[retrieval.CorpusShard] = None
        self._load_index()

    # -- persistence ------------------------------------------------------------------------------------
    def _index_path(self) -> str:
        return os.path.join(self.index_dir, f"{self.index_name}.index.npz")

    def _load_index(self):
        path = self._index_path()
        if os.path.exists(path):
            z = np.load(path, allow_pickle=False)
            emb = torch.from_numpy(z["embedding"].astype(np.float32, copy=False))
            if emb.shape[1] == self.projection_dim:
                self.documents = [str(t) for t in z["content"]]
                self._install(emb.to(self.device))
                print(f"Using existing search index: {self.index_name} ({len(self.documents)} documents)")
                return
        print(f"Creating new search index: {self.index_name}")

    def save_index(self):
        if self.embeddings is None:
            return
        os.makedirs(self.index_dir, exist_ok=True)
        np.savez(self._index_path(), embedding=self.embeddings.cpu().numpy(), content=np.array(self.documents, dtype=str))

    def _install(self, emb: torch.Tensor):
        self.embeddings = emb
        self.shard = retrieval.CorpusShard(emb, id_base=0, precision=self.precision) if emb.shape[0] else None

    # -- the reference's methods ------------------------------------------------------------------------
    def ingest_documents(self, documents: list[str], batch_size: int = 1024, clear_existing: bool = False,
                         persist: bool = False):
        """Encode with the document tower and add to the index (search.py:268-350).  Ids are the positions in the
        index, as the reference's `str(i)` ids."""
        if clear_existing:
            print("Clearing existing documents from index...")
            self.documents, self.embeddings, self.shard = [], None, None
        print(f"Ingesting {len(documents)} documents...")
        with torch.no_grad():
            chunks = [self.model.encode_documents(documents[i: i + batch_size])
                      for i in range(0, len(documents), batch_size)]
        new = torch.cat(chunks, 0).float() if chunks else torch.empty(0, self.projection_dim, device=self.device)
        emb = new if self.embeddings is None else torch.cat([self.embeddings, new], 0)
        self.documents = self.documents + list(documents)
        self._install(emb.contiguous())
        if persist:
            self.save_index()
        print(f"✅ Successfully ingested {len(documents)} documents")

    def search(self, query: str, top_k: int = 10) -> list[dict[str, Any]]:
        """Top-k documents by cosine similarity (search.py:352-402), exact."""
        return self.search_batch([query], top_k)[0]

    def search_batch(self, queries: list[str], top_k: int = 10) -> list[list[dict[str, Any]]]:
        if self.shard is None:
            raise RuntimeError("Search index not initialized")
        with torch.no_grad():
            q = self.model.encode_queries(queries).float()
        print(f"Searching index '{self.index_name}' for top {top_k} results...")
        k = min(top_k, len(self.documents))
        out: list[list[dict[str, Any]]] = []
        if k <= 32:
            s, i = self.shard.search(q, k)
        else:  # the scan keeps at most 32 results per query; deeper lists score every document exactly, then sort
            Qn = ops.l2_normalize_rows(q)
            n = len(self.documents)
            s = torch.empty(len(queries), k, dtype=torch.float32, device=self.device)
            i = torch.empty(len(queries), k, dtype=torch.int64, device=self.device)
            for qi in range(len(queries)):
                sc = torch.from_numpy(ops.candidate_scores(Qn[qi: qi + 1], self.shard.Dn, np.arange(n))).to(self.device)
                order = torch.sort(sc, descending=True, stable=True).indices[:k]  # stable: ties by ascending id
                s[qi], i[qi] = sc[order], order
        scores_all, ids_all = s.cpu().numpy(), i.cpu().numpy()
        for qi in range(len(queries)):
            rows = []
            for sc, di in zip(scores_all[qi], ids_all[qi]):
                if di < 0:
                    continue
                sim = max((1.0 + float(sc)) / 2.0, 0.0)
                rows.append({"id": str(int(di)), "content": self.documents[int(di)], "score": sim, "distance": 1.0 - sim})
            out.append(rows)
        return out

    def get_index_info(self) -> dict[str, Any]:
        n = len(self.documents)
        return {"index_name": self.index_name, "num_docs": n, "indexing_failures": 0,
                "vector_index_sz": round(n * self.projection_dim * 4 / 2 ** 20, 3)}


def build_document_index(max_docs: int = -1, batch_size: int = 1024, model_filename: Optional[str] = None,
                         models_dir: str = MODELS_DIR, projection_dim: Optional[int] = None,
                         redis_url: str = "redis://localhost:6379", redis_host: Optional[str] = None,
                         redis_port: int = 6379, redis_pass: Optional[str] = None, clear_existing: bool = True,
                         **engine_kwargs):
    """Build the document index from the MS MARCO splits (search.py:421-485)."""
    engine = DocumentSearchEngine(model_filename=model_filename, models_dir=models_dir, projection_dim=projection_dim,
                                  redis_url=redis_url, redis_host=redis_host, redis_port=redis_port,
                                  redis_pass=redis_pass, **engine_kwargs)
    if max_docs == -1:
        print("Loading ALL documents from train, validation, and test splits...")
        docs: set[str] = set()
        for split in ["train", "validation", "test"]:
            try:
                split_docs = MSMarcoDataset(split, max_samples=-1).get_unique_passages()
                docs.update(split_docs)
                print(f"    Added {len(split_docs)} documents from {split} split")
            except Exception as e:  # noqa: BLE001
                print(f"    Warning: Could not load {split} split: {e}")
        unique_docs = sorted(docs)
        print(f"Total unique documents across all splits: {len(unique_docs)}")
    else:
        print(f"Loading MS Marco train dataset (max {max_docs} documents)...")
        unique_docs = MSMarcoDataset("train", max_samples=max_docs).get_unique_passages()
        if max_docs > 0 and len(unique_docs) > max_docs:
            unique_docs = random.sample(unique_docs, max_docs)
        print(f"Found {len(unique_docs)} unique documents")
    engine.ingest_documents(unique_docs, batch_size=batch_size, clear_existing=clear_existing, persist=True)
    print(f"Index info: {engine.get_index_info()}")
    return engine


def search_documents(query: str, top_k: int = 10, model_filename: Optional[str] = None, models_dir: str = MODELS_DIR,
                     projection_dim: Optional[int] = None, redis_url: str = "redis://localhost:6379",
                     redis_host: Optional[str] = None, redis_port: Optional[int] = 6379,
                     redis_pass: Optional[str] = None, **engine_kwargs):
    """Search with the auto-selected model and print the results (search.py:488-542)."""
    print(f"Searching for: '{query}'")
    print("-" * 50)
    engine = DocumentSearchEngine(model_filename=model_filename, models_dir=models_dir, projection_dim=projection_dim,
                                  redis_url=redis_url, redis_host=redis_host, redis_port=redis_port,
                                  redis_pass=redis_pass, **engine_kwargs)
    info = engine.get_index_info()
    if info.get("num_docs", 0) == 0:
        print("❌ No documents found in index. Please run with --build-index first (or simultaneously)")
        return None
    print(f"Searching index with {info.get('num_docs', 0)} documents...")
    results = engine.search(query, top_k=top_k)
    print(f"\n🔍 Top {len(results)} results:")
    print("=" * 80)
    for i, r in enumerate(results, 1):
        preview = r["content"][:200] + "..." if len(r["content"]) > 200 else r["content"]
        print(f"{i}. [Score: {r['score']:.4f}] [Distance: {r['distance']:.4f}]")
        print(f"   {preview}")
        if i < len(results):
            print()
    return results
