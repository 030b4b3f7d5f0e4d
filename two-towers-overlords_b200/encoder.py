"""
The reference's actual backbone: the frozen 6-layer MiniLM `BertModel` that `AveragePoolingTower` loads with
`AutoModel.from_pretrained("sentence-transformers/all-MiniLM-L6-v2")` (backend/model.py:24), freezes
(backend/model.py:28-30) and runs under `no_grad` (backend/model.py:51-52); the tower mean-pools `output[0]`, the last
hidden state.

`MiniLMBackbone` keeps the parameter names and shapes of transformers' `BertModel` (a real checkpoint's state dict
loads unchanged; the parameter counts equal `models/e15.lr4.d384.m3_summary.txt:7-9`: 11,918,592 + 10,646,784 +
147,840) and runs the forward pass in libtt_b200.so (`tt_encoder_fwd`: tcgen05 split-bf16 contractions, fp32
LayerNorm / softmax attention kernels).  Without network access the weights are BERT's random init
(`initializer_range` 0.02), which is also what the parity oracle — `BertModel(BertConfig(30522, 384, 6, 12, 1536))`
with the same weights — uses.
"""
from __future__ import annotations

import ctypes

import torch
from torch import Tensor
from torch.nn import Module, ModuleList, Parameter

try:
    from . import _native as N
except ImportError:  # drop-in mode: this directory on sys.path
    import _native as N


class _Cfg:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _w(*shape, std=0.02):
    return Parameter(torch.empty(*shape).normal_(0.0, std), requires_grad=False)


def _zeros(*shape):
    return Parameter(torch.zeros(*shape), requires_grad=False)


def _ones(*shape):
    return Parameter(torch.ones(*shape), requires_grad=False)


class _Dense(Module):
    def __init__(self, n_in, n_out):
        super().__init__()
        self.weight, self.bias = _w(n_out, n_in), _zeros(n_out)


class _LN(Module):
    def __init__(self, n):
        super().__init__()
        self.weight, self.bias = _ones(n), _zeros(n)


class _Emb(Module):
    def __init__(self, n, h):
        super().__init__()
        self.weight = _w(n, h)


class _Embeddings(Module):
    def __init__(self, vocab, max_pos, h):
        super().__init__()
        self.word_embeddings, self.position_embeddings = _Emb(vocab, h), _Emb(max_pos, h)
        self.token_type_embeddings = _Emb(2, h)
        self.LayerNorm = _LN(h)


class _SelfAtt(Module):
    def __init__(self, h):
        super().__init__()
        self.query, self.key, self.value = _Dense(h, h), _Dense(h, h), _Dense(h, h)


class _DenseLN(Module):
    def __init__(self, n_in, n_out):
        super().__init__()
        self.dense, self.LayerNorm = _Dense(n_in, n_out), _LN(n_out)


class _Attention(Module):
    def __init__(self, h):
        super().__init__()
        self.self = _SelfAtt(h)
        self.output = _DenseLN(h, h)


class _Intermediate(Module):
    def __init__(self, h, inter):
        super().__init__()
        self.dense = _Dense(h, inter)


class _Layer(Module):
    def __init__(self, h, inter):
        super().__init__()
        self.attention = _Attention(h)
        self.intermediate = _Intermediate(h, inter)
        self.output = _DenseLN(inter, h)


class _Encoder(Module):
    def __init__(self, h, inter, n):
        super().__init__()
        self.layer = ModuleList(_Layer(h, inter) for _ in range(n))


class _Pooler(Module):
    def __init__(self, h):
        super().__init__()
        self.dense = _Dense(h, h)  # unused by output[0]; kept so that checkpoints load and parameter counts match


class MiniLMBackbone(Module):
    """Frozen BERT encoder (defaults = all-MiniLM-L6-v2: 6 layers, hidden 384, 12 heads, FFN 1536).  `forward`
    returns a 1-tuple holding the last hidden state, like the `output[0]` the reference indexes (model.py:52)."""

    def __init__(self, vocab_size: int = 30522, hidden_size: int = 384, num_hidden_layers: int = 6,
                 num_attention_heads: int = 12, intermediate_size: int = 1536, max_position_embeddings: int = 512,
                 layer_norm_eps: float = 1e-12, max_tokens_per_call: int = 1 << 17):
        super().__init__()
        self.config = _Cfg(vocab_size=vocab_size, hidden_size=hidden_size, num_hidden_layers=num_hidden_layers,
                           num_attention_heads=num_attention_heads, intermediate_size=intermediate_size,
                           max_position_embeddings=max_position_embeddings, layer_norm_eps=layer_norm_eps)
        self.embeddings = _Embeddings(vocab_size, max_position_embeddings, hidden_size)
        self.encoder = _Encoder(hidden_size, intermediate_size, num_hidden_layers)
        self.pooler = _Pooler(hidden_size)
        self.max_tokens_per_call = int(max_tokens_per_call)
        self._prepared = None  # (key, keep-alive tensors, EncoderWeights, layer array)

    @property
    def device(self):
        return self.embeddings.word_embeddings.weight.device

    # -- weights for the C ABI: bf16 (hi, lo) terms of every Linear weight, prepared once (the backbone is frozen) ----
    def _key(self):
        return (str(self.device),) + tuple(p._version for p in self.parameters()) + tuple(p.data_ptr() for p in self.parameters())

    def refresh(self):
        """Call after changing weights by means that do not bump tensor versions; otherwise automatic."""
        self._prepared = None

    def _prepare(self):
        key = self._key()
        if self._prepared is not None and self._prepared[0] == key:
            return self._prepared
        N.ensure_sm100()
        lib = N.load()
        dev, cfg = self.device, self.config
        keep = []

        def f32(t):
            t = t.detach().to(dev, torch.float32).contiguous()
            keep.append(t)
            return t

        def terms(wt):
            wt = f32(wt)
            hi = torch.empty(wt.shape, dtype=torch.bfloat16, device=dev)
            lo = torch.empty_like(hi)
            N.check(lib.tt_split_bf16_terms(N.ptr(wt), wt.shape[0], wt.shape[1], N.ptr(hi), N.ptr(lo), N.stream()),
                    "tt_split_bf16_terms")
            keep.extend((hi, lo))
            return N.ptr(hi), N.ptr(lo)

        layers = (N.EncoderLayer * cfg.num_hidden_layers)()
        for i, ly in enumerate(self.encoder.layer):
            sa, E = ly.attention.self, layers[i]
            E.wqkv_hi, E.wqkv_lo = terms(torch.cat([sa.query.weight, sa.key.weight, sa.value.weight], 0))
            E.bqkv = N.ptr(f32(torch.cat([sa.query.bias, sa.key.bias, sa.value.bias], 0)))
            E.wo_hi, E.wo_lo = terms(ly.attention.output.dense.weight)
            E.bo = N.ptr(f32(ly.attention.output.dense.bias))
            E.ln1_g, E.ln1_b = N.ptr(f32(ly.attention.output.LayerNorm.weight)), N.ptr(f32(ly.attention.output.LayerNorm.bias))
            E.w1_hi, E.w1_lo = terms(ly.intermediate.dense.weight)
            E.b1 = N.ptr(f32(ly.intermediate.dense.bias))
            E.w2_hi, E.w2_lo = terms(ly.output.dense.weight)
            E.b2 = N.ptr(f32(ly.output.dense.bias))
            E.ln2_g, E.ln2_b = N.ptr(f32(ly.output.LayerNorm.weight)), N.ptr(f32(ly.output.LayerNorm.bias))
        W = N.EncoderWeights()
        em = self.embeddings
        W.word_emb, W.pos_emb = N.ptr(f32(em.word_embeddings.weight)), N.ptr(f32(em.position_embeddings.weight))
        W.type_emb = N.ptr(f32(em.token_type_embeddings.weight[0]))
        W.emb_ln_g, W.emb_ln_b = N.ptr(f32(em.LayerNorm.weight)), N.ptr(f32(em.LayerNorm.bias))
        W.vocab, W.max_pos, W.hidden = cfg.vocab_size, cfg.max_position_embeddings, cfg.hidden_size
        W.heads, W.inter, W.n_layers = cfg.num_attention_heads, cfg.intermediate_size, cfg.num_hidden_layers
        W.ln_eps = float(cfg.layer_norm_eps)
        W.layers = ctypes.cast(layers, ctypes.POINTER(N.EncoderLayer))
        torch.cuda.current_stream(dev).synchronize()
        self._prepared = (key, keep, W, layers)
        return self._prepared

    @torch.no_grad()
    def forward(self, input_ids: Tensor, attention_mask: Tensor = None, token_type_ids: Tensor = None,
                check_ids: bool = False, **_):
        if self.device.type != "cuda":
            raise RuntimeError("MiniLMBackbone runs on a CUDA (sm_100a) device only: there is no CPU fallback")
        if token_type_ids is not None and bool((token_type_ids != 0).any()):
            raise NotImplementedError("token_type_ids other than 0 (single-sentence input, model.py:43-45)")
        _, _, W, _ = self._prepare()
        lib, dev = N.load(), self.device
        ids = input_ids.to(dev).contiguous()
        if ids.dtype not in (torch.int64, torch.int32):
            ids = ids.to(torch.int64)
        mask = torch.ones_like(ids) if attention_mask is None else attention_mask.to(dev).contiguous()
        if mask.dtype == torch.bool:
            mask = mask.to(torch.uint8)
        elif mask.dtype not in (torch.int64, torch.int32, torch.uint8):
            mask = mask.to(torch.int64)
        B, L = ids.shape
        H = self.config.hidden_size
        out = torch.empty(B, L, H, dtype=torch.float32, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        step = max(1, self.max_tokens_per_call // max(L, 1))
        ws_bytes = lib.tt_encoder_ws_bytes(min(B, step) * L, H, self.config.intermediate_size)
        ws = N.workspace(ws_bytes, dev)
        for b0 in range(0, B, step):
            b1 = min(B, b0 + step)
            N.check(lib.tt_encoder_fwd(ctypes.byref(W), N.ptr(ids[b0:b1]), N.dtype_code(ids), N.ptr(mask[b0:b1]),
                                       N.dtype_code(mask), b1 - b0, L, N.ptr(out[b0:b1]), N.ptr(err), N.ptr(ws), ws_bytes,
                                       N.stream()), "tt_encoder_fwd")
        if check_ids and int(err.item()) != 0:  # host sync: opt-in (the reference's nn.Embedding raises here)
            raise IndexError("token id out of range of the embedding table")
        return (out,)
