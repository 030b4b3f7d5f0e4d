"""
Two-towers document retrieval model — B200-native drop-in for the reference's backend/model.py.

Same surface as the reference (backend/model.py:13-145): `AveragePoolingTower`, `TwoTowersModel`
(`encode_queries`, `encode_documents`, `encode_documents_batched`, `forward`) and `TripletLoss`, with the
same parameter names (`{query,document}_tower.projection.{0,2}.{weight,bias}`), so backend/training.py,
backend/main.py and backend/search.py run on top of it unchanged.  Underneath, every op is a hand-written
sm_100a kernel in libtt_b200.so (ops.py); there is no eager-PyTorch or CPU fallback.

Backbone: per BASELINE.json's north_star the tower's `pretrained_model(**tokens)[0]` (model.py:51-52) is the
token-embedding lookup E[input_ids] over a 30522x384 table, frozen like the reference
(model.py:28-30) unless `train_table=True` (D2 extension).
Additive: towers also accept pre-tokenised input — a dict with input_ids/attention_mask, an (ids, mask)
tuple, or a TokenBatch — because no tokenizer vocabulary exists offline.
"""
from __future__ import annotations

import os
import zlib
from typing import Optional, Sequence, Tuple, Union

import torch
from torch import Tensor, cat
from torch.nn import Linear, Module, Parameter, ReLU, Sequential

try:
    from . import ops
    from .encoder import MiniLMBackbone
except ImportError:
    import ops
    from encoder import MiniLMBackbone

DEFAULT_MODEL = "sentence-transformers/all-MiniLM-L6-v2"
VOCAB_SIZE = 30522
HIDDEN_SIZE = 384


class TokenBatch:
    """Pre-tokenised texts: ids/mask [B,L] (what the HF tokenizer returns at model.py:43-45)."""

    def __init__(self, input_ids: Tensor, attention_mask: Optional[Tensor] = None):
        self.input_ids = input_ids
        self.attention_mask = attention_mask if attention_mask is not None else torch.ones_like(input_ids)

    def __len__(self):
        return self.input_ids.shape[0]

    def __getitem__(self, sl):
        if isinstance(sl, int):
            sl = slice(sl, sl + 1)
        return TokenBatch(self.input_ids[sl], self.attention_mask[sl])

    def to(self, device, non_blocking=True):
        return TokenBatch(self.input_ids.to(device, non_blocking=non_blocking),
                          self.attention_mask.to(device, non_blocking=non_blocking))


TextsOrTokens = Union[Sequence[str], TokenBatch, dict, Tuple[Tensor, Tensor]]


class HashTokenizer:
    """Offline stand-in for the WordPiece tokenizer (no vocab.txt exists without network): lower-cased
    whitespace/punctuation split, each word hashed (CRC32) into [999, vocab), [CLS]=101 ... [SEP]=102,
    pad id 0.  Same call signature and return keys as the HF tokenizer call at model.py:43-45."""

    def __init__(self, vocab_size: int = VOCAB_SIZE):
        self.vocab_size = vocab_size

    def _encode(self, text: str, max_length: int) -> list[int]:
        words = "".join(c if c.isalnum() else " " for c in text.lower()).split()
        lo = min(999, self.vocab_size // 2)
        body = [lo + zlib.crc32(w.encode()) % (self.vocab_size - lo) for w in words][: max_length - 2]
        return [101] + body + [102]

    def __call__(self, texts, padding=True, truncation=True, return_tensors="pt", max_length=512):
        rows = [self._encode(t, max_length) for t in texts]
        L = max((len(r) for r in rows), default=1)
        ids = torch.zeros(len(rows), L, dtype=torch.int64)
        mask = torch.zeros(len(rows), L, dtype=torch.int64)
        for i, r in enumerate(rows):
            ids[i, : len(r)] = torch.tensor(r, dtype=torch.int64)
            mask[i, : len(r)] = 1
        return {"input_ids": ids, "token_type_ids": torch.zeros_like(ids), "attention_mask": mask}


def _load_tokenizer(model_name: str, vocab_size: int):
    """HF tokenizer when its files are in the local cache, else the offline HashTokenizer."""
    cache = os.path.join(os.environ.get("HF_HOME", os.path.expanduser("~/.cache/huggingface")), "hub",
                         "models--" + model_name.replace("/", "--"))
    if os.environ.get("TT_TOKENIZER", "auto") != "hash" and os.path.isdir(cache):
        try:
            from transformers import AutoTokenizer

            return AutoTokenizer.from_pretrained(model_name, local_files_only=True)
        except Exception:  # noqa: BLE001 - any cache problem means: no usable tokenizer offline
            pass
    return HashTokenizer(vocab_size)


class _Config:
    def __init__(self, hidden_size: int, vocab_size: int):
        self.hidden_size = hidden_size
        self.vocab_size = vocab_size


class _WordEmbeddings(Module):
    def __init__(self, vocab_size: int, hidden_size: int, dtype: torch.dtype):
        super().__init__()
        # nn.Embedding default init N(0,1); random-init stands in for the MiniLM checkpoint (no network)
        self.weight = Parameter(torch.empty(vocab_size, hidden_size).normal_().to(dtype), requires_grad=False)


class _Embeddings(Module):
    def __init__(self, vocab_size, hidden_size, dtype):
        super().__init__()
        self.word_embeddings = _WordEmbeddings(vocab_size, hidden_size, dtype)


class TokenTableBackbone(Module):
    """`pretrained_model` of the tower: exposes .config.hidden_size and .device like the HF model the
    reference holds (model.py:24-26,48); its only weight keeps the HF key
    `embeddings.word_embeddings.weight`, so a reference checkpoint's table loads with strict=False."""

    def __init__(self, vocab_size: int = VOCAB_SIZE, hidden_size: int = HIDDEN_SIZE, dtype=torch.float32):
        super().__init__()
        self.config = _Config(hidden_size, vocab_size)
        self.embeddings = _Embeddings(vocab_size, hidden_size, dtype)

    @property
    def table(self) -> Parameter:
        return self.embeddings.word_embeddings.weight

    @property
    def device(self):
        return self.table.device


class AveragePoolingTower(Module):
    """Token-row gather + masked mean + L2 normalise + trainable 2-layer projection (model.py:13-72)."""

    def __init__(
        self,
        model_name: str = DEFAULT_MODEL,
        projection_dim: int = 128,
        vocab_size: int = VOCAB_SIZE,
        hidden_size: int = HIDDEN_SIZE,
        table_dtype: torch.dtype = torch.float32,
        train_table: bool = False,
        precision: Optional[str] = None,
        backbone: str = "table",
    ):
        super().__init__()
        self.tokenizer = _load_tokenizer(model_name, vocab_size)
        # "table": the north_star's backbone, E[input_ids].  "minilm": the reference's actual one, the
        # frozen 6-layer MiniLM BertModel of model.py:24 (encoder.MiniLMBackbone, forward on the device)
        if backbone not in ("table", "minilm"):
            raise ValueError(f"backbone must be 'table' or 'minilm', got {backbone!r}")
        self.backbone = backbone
        if backbone == "minilm":
            if train_table:
                raise ValueError("the MiniLM backbone is frozen (model.py:28-30); train_table needs backbone='table'")
            self.pretrained_model = MiniLMBackbone(vocab_size=vocab_size, hidden_size=hidden_size)
        else:
            self.pretrained_model = TokenTableBackbone(vocab_size, hidden_size, table_dtype)
        self.embedding_dim = self.pretrained_model.config.hidden_size
        # frozen backbone (model.py:28-30) unless the D2 extension is requested
        for param in self.pretrained_model.parameters():
            param.requires_grad = bool(train_table)
        self.projection = Sequential(
            Linear(self.embedding_dim, projection_dim),
            ReLU(),
            Linear(projection_dim, projection_dim),
        )
        self.precision = precision or ops.default_precision(self.embedding_dim, projection_dim)

    # -- tokens ------------------------------------------------------------------------------------
    def tokenize(self, texts: TextsOrTokens) -> Tuple[Tensor, Tensor]:
        if isinstance(texts, TokenBatch):
            return texts.input_ids, texts.attention_mask
        if isinstance(texts, dict) or hasattr(texts, "keys"):
            return texts["input_ids"], texts.get("attention_mask")
        if isinstance(texts, tuple) and len(texts) == 2 and isinstance(texts[0], Tensor):
            return texts
        if len(texts) == 0:
            raise ValueError("cannot encode an empty batch of texts")
        tokens = self.tokenizer(list(texts), padding=True, truncation=True, return_tensors="pt", max_length=512)
        return tokens["input_ids"], tokens["attention_mask"]

    def pooled(self, texts: TextsOrTokens) -> Tensor:
        """Normalised mean-pooled embeddings [B, hidden] (model.py:48-56)."""
        ids, mask = self.tokenize(texts)
        if self.backbone == "minilm":
            # model.py:51-56: last hidden state -> masked mean -> normalise.  The pool kernel gathers rows of a
            # "table" by index: here the table is the hidden states and the index of token (b, l) is b * L + l.
            if mask is None:
                mask = torch.ones_like(ids)
            hidden = self.pretrained_model(input_ids=ids, attention_mask=mask)[0]
            B, L, H = hidden.shape
            pos = torch.arange(B * L, device=hidden.device, dtype=torch.int64).view(B, L)
            return ops.pool(hidden.view(B * L, H), pos, mask.to(hidden.device))
        return ops.pool(self.pretrained_model.table, ids, mask)

    def forward(self, texts: TextsOrTokens) -> Tensor:
        x = self.pooled(texts)
        p0, p2 = self.projection[0], self.projection[2]
        return ops.mlp(x, p0.weight, p0.bias, p2.weight, p2.bias, self.precision)


class TwoTowersModel(Module):
    """Two-towers architecture for document retrieval (model.py:75-121)."""

    def __init__(self, projection_dim: int = 128, model_name: str = DEFAULT_MODEL, **tower_kwargs):
        super().__init__()
        self.query_tower = AveragePoolingTower(model_name, projection_dim, **tower_kwargs)
        self.document_tower = AveragePoolingTower(model_name, projection_dim, **tower_kwargs)

    def encode_queries(self, queries: TextsOrTokens) -> Tensor:
        return self.query_tower(queries)

    def encode_documents(self, documents: TextsOrTokens) -> Tensor:
        return self.document_tower(documents)

    def encode_documents_batched(self, documents: TextsOrTokens, batch_size: int = 1024) -> Tensor:
        """Chunked encode of a large collection (model.py:95-112).  Chunks stay on the device: the
        reference's CPU round trip only existed to spare a 24 GB card; 180 GB of HBM hold 8.8 M x 384."""
        n = len(documents)
        if n <= batch_size:
            return self.encode_documents(documents)
        chunks = [self.encode_documents(documents[i: i + batch_size]) for i in range(0, n, batch_size)]
        return cat(chunks, dim=0)

    def forward(self, queries: TextsOrTokens, documents: TextsOrTokens) -> Tuple[Tensor, Tensor]:
        return self.encode_queries(queries), self.encode_documents(documents)

    def projection_parameters(self) -> list[Parameter]:
        """The 8 trainable tensors in C-ABI order: Wq1,bq1,Wq2,bq2,Wd1,bd1,Wd2,bd2."""
        out = []
        for tower in (self.query_tower, self.document_tower):
            out += [tower.projection[0].weight, tower.projection[0].bias,
                    tower.projection[2].weight, tower.projection[2].bias]
        return out


class TripletLoss(Module):
    """Triplet loss with cosine distance (model.py:125-145)."""

    def __init__(self, margin: float = 0.3):
        super().__init__()
        self.margin = margin

    def cosine_distance(self, x: Tensor, y: Tensor) -> Tensor:
        return 1 - torch.nn.functional.cosine_similarity(x, y, dim=1)

    def forward(self, anchor: Tensor, positive: Tensor, negative: Tensor) -> Tensor:
        return ops.triplet_loss(anchor, positive, negative, self.margin)
