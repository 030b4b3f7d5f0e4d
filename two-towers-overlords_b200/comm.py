"""
Peer-memory data parallelism for one NVLink/NVSwitch box.  The reference trains on one device
(training.py:47-51: `loss.backward(); optimizer.step()`); a data-parallel run needs exactly one exchange per step,
the sum of the flat projection gradient.  Here that exchange and the optimiser are ONE kernel
(`tt_dp_reduce_adam`: reduce-scatter over peer stores -> Adam on the local slice -> all-gather of the new
parameters into every rank's memory), not an NCCL call followed by an Adam launch.

torch.distributed is only the rendezvous: it ships the 64-byte CUDA IPC handles once.
"""
from __future__ import annotations

import ctypes
from ctypes import c_void_p
from typing import List, Optional

import torch

try:
    from . import _native as N
except ImportError:
    import _native as N

HANDLE_BYTES = 64
_TYPESTR = {torch.float32: "<f4", torch.int32: "<i4", torch.int64: "<i8", torch.uint8: "|u1", torch.float64: "<f8"}


class _RawView:
    """__cuda_array_interface__ carrier so torch can wrap library-owned device memory without copying."""

    def __init__(self, ptr: int, shape, dtype, owner):
        self.owner = owner  # keeps the segment alive as long as the tensor lives
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": _TYPESTR[dtype], "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerSegment:
    """Device memory other processes on the box can map (cudaMalloc + CUDA IPC handle), zero-filled."""

    def __init__(self, nbytes: int, device):
        N.ensure_sm100()
        self.device = torch.device(device)
        self.nbytes = int(nbytes)
        ptr = c_void_p()
        handle = ctypes.create_string_buffer(HANDLE_BYTES)
        with torch.cuda.device(self.device):
            N.check(N.load().tt_peer_alloc(self.nbytes, ctypes.byref(ptr), handle), "tt_peer_alloc")
        self.ptr = int(ptr.value)
        self.handle = bytes(handle.raw)
        self._mapped: List[int] = []

    def tensor(self, offset_bytes: int, shape, dtype) -> torch.Tensor:
        assert 0 <= offset_bytes < self.nbytes
        return torch.as_tensor(_RawView(self.ptr + offset_bytes, shape, dtype, self), device=self.device)

    def open_peer(self, handle: bytes) -> int:
        out = c_void_p()
        with torch.cuda.device(self.device):
            N.check(N.load().tt_peer_open(ctypes.c_char_p(handle), ctypes.byref(out)), "tt_peer_open")
        self._mapped.append(int(out.value))
        return int(out.value)

    def close(self):
        lib = N.load()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for p in self._mapped:
                lib.tt_peer_close(c_void_p(p))
            self._mapped = []
            if self.ptr:
                lib.tt_peer_free(c_void_p(self.ptr))
                self.ptr = 0


def exchange_handles(handle: bytes, world: int, group=None) -> List[bytes]:
    out: List[Optional[bytes]] = [None] * world
    torch.distributed.all_gather_object(out, handle, group=group)
    return [bytes(h) for h in out]


class DpExchange:
    """One rank's end of the fused gradient exchange.  `flat_p` ([n_param] fp32) lives in this rank's peer segment:
    the trainer aliases its nn.Parameters to it and peers store new parameter slices straight into it;
    `loss` ([1]) is slot n_param of the same segment = the global loss of the last exchanged step."""

    def __init__(self, n_param: int, world: int, rank: int, device, segments: Optional[List[int]] = None,
                 own: Optional[PeerSegment] = None, group=None):
        lib = N.load()
        self.n, self.world, self.rank, self.device = int(n_param), int(world), int(rank), torch.device(device)
        self.seg_bytes = int(lib.tt_dp_segment_bytes(self.n, self.world))
        if self.seg_bytes == 0:
            raise N.NativeError(f"peer exchange supports 1..8 ranks, got {world}")
        self.own = own if own is not None else PeerSegment(self.seg_bytes, self.device)
        if segments is None:  # real multi-process setup: ship IPC handles through torch.distributed
            if self.world > 1:
                handles = exchange_handles(self.own.handle, self.world, group)
                segments = [self.own.ptr if r == self.rank else self.own.open_peer(handles[r])
                            for r in range(self.world)]
                torch.distributed.barrier(group=group)
            else:
                segments = [self.own.ptr]
        self.segments = (c_void_p * self.world)(*[c_void_p(int(s)) for s in segments])
        self.flat_p = self.own.tensor(0, (self.n,), torch.float32)
        self.loss = self.own.tensor(self.n * 4, (1,), torch.float32)
        self.state = torch.zeros(4, dtype=torch.float64, device=self.device)
        self.ctl = torch.zeros(16, dtype=torch.int32, device=self.device)

    @classmethod
    def virtual_ranks(cls, n_param: int, world: int, device) -> List["DpExchange"]:
        """`world` ranks inside ONE process on ONE device (tests): the same kernel and protocol, segments are plain
        device pointers instead of IPC mappings; the caller must launch every rank on its own stream."""
        lib = N.load()
        nbytes = int(lib.tt_dp_segment_bytes(n_param, world))
        owns = [PeerSegment(nbytes, device) for _ in range(world)]
        ptrs = [o.ptr for o in owns]
        return [cls(n_param, world, r, device, segments=ptrs, own=owns[r]) for r in range(world)]

    def reduce_adam(self, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, lr: float,
                    betas=(0.9, 0.999), eps: float = 1e-8, max_ctas: int = 0):
        """grad [n_param + 1] (last slot: this rank's loss term).  Asynchronous on the current stream."""
        assert grad.numel() == self.n + 1 and grad.dtype == torch.float32
        assert exp_avg.numel() == self.n and exp_avg_sq.numel() == self.n
        N.check(N.load().tt_dp_reduce_adam(self.segments, self.world, self.rank, self.n, N.ptr(grad), N.ptr(exp_avg),
                                           N.ptr(exp_avg_sq), float(lr), float(betas[0]), float(betas[1]), float(eps),
                                           N.ptr(self.state), N.ptr(self.ctl), int(max_ctas), N.stream()),
                "tt_dp_reduce_adam")

    def check(self):
        """Host sync: raises if a peer failed to answer inside the kernel's timeout."""
        if int(self.ctl[3].item()) != 0:
            raise N.NativeError("peer exchange timed out waiting for another rank")

    def reset_timing(self):
        self.ctl[4:].zero_()

    def timing(self):
        """(mean us until all ranks' gradient slices landed, mean us from there to kernel end, calls, mean us of the own push) — CTA 0's view."""
        d = self.ctl[4:12].view(torch.int64).tolist()
        n = max(1, d[2])
        return d[0] / n / 1e3, d[1] / n / 1e3, d[2], d[3] / n / 1e3

    def close(self):
        self.own.close()


class PeerLists:
    """Per-shard top-k lists of the corpus scan kept in peer memory: every rank writes its [Q,k]
    (score f32, id i64) list into its own segment, then `merge()` = barrier -> `tt_peer_topk_merge` (each rank reads
    all lists in place over NVLink and merges by (score desc, id asc)) -> barrier.  Replaces the all-gather + merge."""

    def __init__(self, Q: int, k: int, world: int, rank: int, device, segments: Optional[List[int]] = None,
                 own: Optional[PeerSegment] = None, group=None):
        self.Q, self.k, self.world, self.rank, self.device = int(Q), int(k), int(world), int(rank), torch.device(device)
        self.off_id = (self.Q * self.k * 4 + 255) // 256 * 256
        self.off_flag = self.off_id + (self.Q * self.k * 8 + 255) // 256 * 256
        self.nbytes = self.off_flag + 256
        self.own = own if own is not None else PeerSegment(self.nbytes, self.device)
        if segments is None:
            if self.world > 1:
                handles = exchange_handles(self.own.handle, self.world, group)
                segments = [self.own.ptr if r == self.rank else self.own.open_peer(handles[r]) for r in range(self.world)]
                torch.distributed.barrier(group=group)
            else:
                segments = [self.own.ptr]
        arr = lambda off: (c_void_p * self.world)(*[c_void_p(int(s_) + off) for s_ in segments])  # noqa: E731
        self.p_score, self.p_id, self.p_flag = arr(0), arr(self.off_id), arr(self.off_flag)
        self.score = self.own.tensor(0, (self.Q, self.k), torch.float32)
        self.ids = self.own.tensor(self.off_id, (self.Q, self.k), torch.int64)
        self.ctl = torch.zeros(4, dtype=torch.int32, device=self.device)

    @classmethod
    def virtual_ranks(cls, Q: int, k: int, world: int, device) -> List["PeerLists"]:
        probe = cls.__new__(cls)
        off_id = (Q * k * 4 + 255) // 256 * 256
        nbytes = off_id + (Q * k * 8 + 255) // 256 * 256 + 256
        owns = [PeerSegment(nbytes, device) for _ in range(world)]
        ptrs = [o.ptr for o in owns]
        del probe
        return [cls(Q, k, world, r, device, segments=ptrs, own=owns[r]) for r in range(world)]

    def barrier(self):
        N.check(N.load().tt_peer_barrier(self.p_flag, self.world, self.rank, N.ptr(self.ctl), N.stream()),
                "tt_peer_barrier")

    def merge(self, top_s: torch.Tensor, top_i: torch.Tensor):
        """This rank's shard list [Q,k] -> the global [Q,k] (same on every rank).  Asynchronous on the current stream."""
        self.score.copy_(top_s)
        self.ids.copy_(top_i)
        out_s = torch.empty(self.Q, self.k, dtype=torch.float32, device=self.device)
        out_i = torch.empty(self.Q, self.k, dtype=torch.int64, device=self.device)
        self.barrier()  # every list complete
        N.check(N.load().tt_peer_topk_merge(self.p_score, self.p_id, self.world, self.Q, self.k, N.ptr(out_s),
                                            N.ptr(out_i), N.stream()), "tt_peer_topk_merge")
        self.barrier()  # nobody overwrites a list that a peer is still reading
        return out_s, out_i

    def check(self):
        if int(self.ctl[3].item()) != 0:
            raise N.NativeError("peer barrier timed out waiting for another rank")

    def close(self):
        self.own.close()
