"""FSDP2 collectives over NVLink symmetric memory (what bench.py runs at N > 1; NCCL stays the library default).

FSDP2 (ref: touchnet/models/helper_func.py:134-202) lets a module group swap its communication primitives
(`FSDPModule.set_custom_reduce_scatter / set_custom_all_gather`, torch/distributed/fsdp/_fully_shard/_fsdp_api.py).
`install(model, group, device, mode=...)` gives every group buffers in torch symmetric memory (every rank maps every rank's
buffer, CUDA IPC over NVLink / NVSwitch) and one of two forms:

  mode="push" (measured best, profiles/r02_bench_n{2,4,8}_*.json)
    all-gather      every rank copies its shard into every peer's output buffer with plain device-to-device copies
                    (copy engines, no SM), one device-side barrier                                      -> PushAllGather
    reduce-scatter  direct: FSDP2's chunk_cat copy-in (a kernel on the COMPUTE stream) is bypassed; chunk p of every
                    unsharded gradient (bf16 under the mixed-precision policy) is pushed from where autograd left it into
                    peer p's receive slot, one barrier, then tn_reduce_bf16_to_f32 adds the N chunks in rank order in fp32
                    (= fp32 reduce of the bf16 gradients, deterministic).  staged: FSDP2's [world, shard] fp32 input buffer
                    is filled as usual and its chunks are pushed                                          -> PushReduceScatter
  mode="pull"  (round 1's idea; measured SLOWER than NCCL: too few bytes in flight per CTA)
    tn_peer_reduce_scatter_f32 / tn_peer_all_gather read the peers' buffers                 -> PeerReduceScatter / PeerAllGather

Why: NCCL's ring kernels share SMs / L2 / HBM with the backward GEMMs they overlap (+40 ms of a 408 ms step at N=2) and
FSDP2's copy-in runs on the compute stream (+26 ms).  With pushes the tensor-core kernels keep the whole chip: 408 -> 364 ms
per step at N=2, 425 -> 373 ms at N=8 (154 k -> 175 k tokens/s).

Ordering: a buffer of the ring of three is rewritten only after later barriers which every reader reaches after it has
consumed the buffer (stream order); all-gather and reduce-scatter barriers use different signal-pad channels because the
two collectives overlap in backward.  Parity on hardware: tools/check_fsdp.py (TN_FSDP_PEER=1|push, TN_FSDP_DIRECT=0|1) on
2 GPUs, the rank-0 loss of every bench line at N = 2, 4, 8; wiring on CPU: tests/test_parallel_gloo.py (gloo, shared-memory
files standing in for symmetric memory, torch standing in for the kernels).
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional, Sequence

import torch
import torch.distributed as dist
from torch.distributed.fsdp._fully_shard._fsdp_api import AllGather, ReduceScatter

from . import _lib

RING = 3        # FSDP2 keeps at most two buffers of a kind alive (current + prefetched); one spare


def _launch_reduce_scatter(ptrs: list, shard_offset: int, out: torch.Tensor, numel: int, scale: float, max_ctas: int) -> None:
    arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
    _lib.call("tn_peer_reduce_scatter_f32", arr, len(ptrs), shard_offset, out.data_ptr(), numel, float(scale), max_ctas,
              torch.cuda.current_stream().cuda_stream)


def _launch_all_gather(ptrs: list, bytes_each: int, out: torch.Tensor, max_ctas: int) -> None:
    arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
    _lib.call("tn_peer_all_gather", arr, len(ptrs), bytes_each, out.data_ptr(), max_ctas,
              torch.cuda.current_stream().cuda_stream)


def _launch_reduce_bf16(ptrs: list, out: torch.Tensor, numel: int, scale: float, max_ctas: int) -> None:
    arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
    _lib.call("tn_reduce_bf16_to_f32", arr, len(ptrs), out.data_ptr(), numel, float(scale), max_ctas,
              torch.cuda.current_stream().cuda_stream)


# ---- FSDP2's reduce-scatter copy-in, bypassed for the direct-push reduce-scatter -------------------------------------------
# torch.distributed.fsdp._fully_shard._fsdp_collectives.foreach_reduce copies (and casts) every unsharded gradient into one
# [world, shard] buffer with a chunk_cat kernel ON THE COMPUTE STREAM (26 ms of a 407 ms step at N=2) before it calls the
# reduce-scatter.  A chunk of a dim-0-sharded gradient is contiguous, so the direct form needs no staging at all: the
# reduce-scatter pushes each chunk from where autograd left it.  The hook below stands in for the copy-in when (and only
# when) the buffer belongs to an installed PushReduceScatter and every gradient qualifies; anything else takes torch's path.
_DIRECT: list = []          # installed PushReduceScatter objects
_orig_copy_in = None


def _copy_in_hook(unsharded_grads, reduce_scatter_input, world_size):
    ptr = reduce_scatter_input.data_ptr()
    for comm in _DIRECT:
        if comm._expect_input == ptr and comm.can_direct(unsharded_grads, world_size):
            comm._stash = (ptr, list(unsharded_grads))       # keeps the gradients alive until their pushes are enqueued
            return
    return _orig_copy_in(unsharded_grads, reduce_scatter_input, world_size)


def _patch_copy_in() -> bool:
    global _orig_copy_in
    if _orig_copy_in is not None:
        return True
    try:
        from torch.distributed.fsdp._fully_shard import _fsdp_collectives as fc
        _orig_copy_in = fc.foreach_reduce_scatter_copy_in
        fc.foreach_reduce_scatter_copy_in = _copy_in_hook
        return True
    except Exception:        # a torch without this seam: the staged path (copy-in + push of the staged chunks) still works
        _orig_copy_in = None
        return False


def _barrier(mem, channel: int) -> None:
    """all-gather barriers on signal-pad channel 0, reduce-scatter barriers on channel 1: FSDP2 runs the two on different
    streams and they overlap in backward (prefetched all-gather of block i-1 next to the reduce-scatter of block i)."""
    try:
        mem.barrier(channel)
    except TypeError:            # test doubles without channels
        mem.barrier()


class _PeerPool:
    """Symmetric communication buffers of one process group, handed out FSDP2-style through `allocate`."""

    def __init__(self, group: dist.ProcessGroup, device, mem=None):
        from .tensor_parallel import SymmPeerMemory
        self.group = group
        self.size, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.mem = mem if mem is not None else SymmPeerMemory(group, torch.device(device))
        self._rings: dict = {}       # (numel, dtype) -> [list of per-rank views] * RING
        self._next: dict = {}
        self._by_ptr: dict = {}      # data_ptr of this rank's buffer -> per-rank views

    def allocate(self, size: Sequence[int], *, dtype: torch.dtype, device) -> torch.Tensor:
        numel = math.prod(int(s) for s in size)
        key = (numel, dtype)
        ring = self._rings.setdefault(key, [])
        i = self._next.get(key, 0)
        if i >= len(ring):
            if len(ring) < RING:
                views = self.mem.alloc((numel,), dtype)          # collective: every rank allocates in the same order
                ring.append(views)
                self._by_ptr[views[self.rank].data_ptr()] = views
            else:
                i = 0
        self._next[key] = (i + 1) % RING if len(ring) == RING else i + 1
        return ring[i][self.rank].view(*[int(s) for s in size])

    def views_of(self, t: torch.Tensor) -> list:
        v = self._by_ptr.get(t.data_ptr())
        if v is None:
            raise _lib.TouchNetB200Error("peer collective called on a buffer that did not come from its allocate()")
        return v


class PeerReduceScatter(ReduceScatter):
    """fp32 gradient reduce-scatter of one FSDP2 group as a pull over peer memory (SUM or AVG)."""

    def __init__(self, pool: _PeerPool, max_ctas: int = 32):
        self.pool, self.max_ctas = pool, max_ctas

    def allocate(self, size, *, dtype, device) -> torch.Tensor:
        return self.pool.allocate(size, dtype=dtype, device=device)

    def __call__(self, output_tensor, input_tensor, group, op, async_op: bool = False):
        if input_tensor.dtype != torch.float32:
            raise _lib.TouchNetB200Error("PeerReduceScatter handles fp32 gradients (reduce_dtype=float32, the reference's default)")
        pool = self.pool
        n = output_tensor.numel()
        if input_tensor.numel() != n * pool.size:
            raise _lib.TouchNetB200Error("reduce-scatter input must be world_size x output")
        if op == dist.ReduceOp.AVG:
            scale = 1.0 / pool.size
        elif op == dist.ReduceOp.SUM:
            scale = 1.0
        else:
            raise _lib.TouchNetB200Error(f"PeerReduceScatter: unsupported reduce op {op}")
        views = pool.views_of(input_tensor)
        _barrier(pool.mem, 1)                                    # every rank's copy-in of this group has completed
        _launch_reduce_scatter([v.data_ptr() for v in views], pool.rank * n, output_tensor, n, scale, self.max_ctas)
        return None                                              # stream-ordered on the caller's (reduce-scatter) stream


class PeerAllGather(AllGather):
    """bf16 (or fp32) parameter all-gather of one FSDP2 group as a pull over peer memory.  FSDP2 gathers in place: the
    input is this rank's slice of the output buffer, so every rank's shard already sits in its own symmetric buffer."""

    def __init__(self, pool: _PeerPool, max_ctas: int = 32):
        self.pool, self.max_ctas = pool, max_ctas

    def allocate(self, size, *, dtype, device) -> torch.Tensor:
        return self.pool.allocate(size, dtype=dtype, device=device)

    def __call__(self, output_tensor, input_tensor, group, async_op: bool = False):
        pool = self.pool
        n = input_tensor.numel()
        es = input_tensor.element_size()
        if output_tensor.numel() != n * pool.size or (n * es) % 16 != 0:
            raise _lib.TouchNetB200Error("all-gather output must be world_size x input, shard bytes a multiple of 16")
        views = pool.views_of(output_tensor)
        if input_tensor.data_ptr() != output_tensor.data_ptr() + pool.rank * n * es:
            output_tensor.view(-1)[pool.rank * n:(pool.rank + 1) * n].copy_(input_tensor.reshape(-1))
        _barrier(pool.mem, 0)                                    # every rank's shard is in its own buffer
        ptrs = [views[p].data_ptr() + p * n * es for p in range(pool.size)]
        _launch_all_gather(ptrs, n * es, output_tensor, self.max_ctas)
        return None


class PushAllGather(AllGather):
    """Parameter all-gather as COPY-ENGINE pushes: every rank writes its shard into every peer's (symmetric) output buffer
    with plain device-to-device copies over NVLink - no SM is involved, so the tensor-core kernels the gather overlaps
    with keep the whole chip (measured at N=2: NCCL's ring kernels slow the GEMMs they share the SMs / L2 with by 14 %;
    the pull kernels above were slower still, profiles/README.md) - then one device-side barrier."""

    def __init__(self, pool: _PeerPool):
        self.pool = pool

    def allocate(self, size, *, dtype, device) -> torch.Tensor:
        return self.pool.allocate(size, dtype=dtype, device=device)

    def __call__(self, output_tensor, input_tensor, group, async_op: bool = False):
        pool = self.pool
        n = input_tensor.numel()
        if output_tensor.numel() != n * pool.size:
            raise _lib.TouchNetB200Error("all-gather output must be world_size x input")
        views = pool.views_of(output_tensor)
        src = input_tensor.reshape(-1)
        in_place = input_tensor.data_ptr() == output_tensor.data_ptr() + pool.rank * n * input_tensor.element_size()
        for k in range(pool.size):                       # start with the next rank: the pushes of different ranks fan out
            p = (pool.rank + 1 + k) % pool.size
            if p == pool.rank and in_place:
                continue
            views[p].view(-1)[pool.rank * n:(pool.rank + 1) * n].copy_(src, non_blocking=True)
        _barrier(pool.mem, 0)                            # every rank's pushes into this buffer have landed
        return None


class PushReduceScatter(ReduceScatter):
    """Gradient reduce-scatter (fp32 result): copy-engine pushes of every peer's chunk into that peer's receive slots, one
    barrier, then ONE local kernel adds the world_size chunks in rank order (deterministic, HBM-bound, a few CTAs).

    Direct form (default when FSDP2's copy-in seam can be patched, `direct=True`): the chunks are pushed straight from the
    unsharded gradients autograd produced - bf16 under the reference's mixed-precision policy, so half the NVLink bytes -
    and summed in fp32 by tn_reduce_bf16_to_f32; FSDP2's chunk_cat copy-in (a kernel on the compute stream) is skipped.
    Staged form: FSDP2 copies into the [world, shard] input buffer as usual and the chunks are pushed from there."""

    def __init__(self, pool: _PeerPool, max_ctas: int = 32, direct: bool = True):
        self.pool, self.max_ctas = pool, max_ctas
        self._recv: dict = {}       # (numel, dtype) -> ring of per-rank receive buffers [world * shard]
        self._next: dict = {}
        self._expect_input = None   # data_ptr of the reduce-scatter input buffer handed out last
        self._stash = None          # (input data_ptr, unsharded gradients) left by the copy-in hook
        self.direct = bool(direct) and _patch_copy_in()
        if self.direct:
            _DIRECT.append(self)

    def allocate(self, size, *, dtype, device) -> torch.Tensor:
        t = torch.empty(*[int(s) for s in size], dtype=dtype, device=device)        # peers never read this buffer
        self._expect_input = t.data_ptr()
        return t

    def can_direct(self, grads, world_size: int) -> bool:
        if world_size != self.pool.size or world_size < 2 or not grads:
            return False        # (a 1-rank group: FSDP2 copies the input buffer itself instead of calling the reduce-scatter)
        dt = grads[0].dtype
        if dt not in (torch.bfloat16, torch.float32):
            return False
        for g in grads:
            if g.dtype != dt or not g.is_contiguous() or g.dim() == 0 or g.shape[0] % world_size != 0:
                return False
            if (g.numel() // world_size * g.element_size()) % 16 != 0:
                return False
        return True

    def _slot(self, numel: int, dtype):
        key = (numel, dtype)
        ring = self._recv.setdefault(key, [])
        i = self._next.get(key, 0)
        if i >= len(ring):
            ring.append(self.pool.mem.alloc((numel,), dtype))                        # collective, same order on all ranks
        self._next[key] = (i + 1) % RING
        return ring[i]

    def __call__(self, output_tensor, input_tensor, group, op, async_op: bool = False):
        pool = self.pool
        n = output_tensor.numel()
        if input_tensor.numel() != n * pool.size:
            raise _lib.TouchNetB200Error("reduce-scatter input must be world_size x output")
        if op == dist.ReduceOp.AVG:
            scale = 1.0 / pool.size
        elif op == dist.ReduceOp.SUM:
            scale = 1.0
        else:
            raise _lib.TouchNetB200Error(f"PushReduceScatter: unsupported reduce op {op}")
        stash, self._stash = self._stash, None
        if stash is not None and stash[0] == input_tensor.data_ptr():
            # ---- direct: chunk p of every gradient -> slot [rank] of peer p's receive buffer (own chunk: a local copy) ----
            grads = stash[1]
            dt = grads[0].dtype
            recv = self._slot(n * pool.size, dt)
            cur = torch.cuda.current_stream() if grads[0].is_cuda else None
            off = 0
            for g in grads:
                flat = g.reshape(-1)
                c = flat.numel() // pool.size
                for k in range(pool.size):
                    p = (pool.rank + k) % pool.size
                    recv[p].view(pool.size, n)[pool.rank, off:off + c].copy_(flat[p * c:(p + 1) * c], non_blocking=True)
                if cur is not None:
                    g.record_stream(cur)         # allocated on the compute stream, read by copies on this (reduce-scatter) stream
                off += c
            if off != n:
                raise _lib.TouchNetB200Error("direct reduce-scatter: gradient chunks do not add up to the shard size")
            _barrier(pool.mem, 1)
            mine = recv[pool.rank]
            es = mine.element_size()
            ptrs = [mine.data_ptr() + q * n * es for q in range(pool.size)]
            if output_tensor.dtype != torch.float32:
                raise _lib.TouchNetB200Error("PushReduceScatter produces fp32 (reduce_dtype=float32, the reference's default)")
            if dt == torch.bfloat16:
                _launch_reduce_bf16(ptrs, output_tensor, n, scale, self.max_ctas)
            else:
                _launch_reduce_scatter(ptrs, 0, output_tensor, n, scale, self.max_ctas)
            return None
        # ---- staged: FSDP2's copy-in filled `input_tensor` ([world, shard], reduce dtype) ----
        if input_tensor.dtype != torch.float32:
            raise _lib.TouchNetB200Error("PushReduceScatter handles fp32 reduce buffers (reduce_dtype=float32, the reference's default)")
        recv = self._slot(n * pool.size, torch.float32)
        flat = input_tensor.reshape(-1)
        for k in range(1, pool.size):
            p = (pool.rank + k) % pool.size
            recv[p].view(-1)[pool.rank * n:(pool.rank + 1) * n].copy_(flat[p * n:(p + 1) * n], non_blocking=True)
        _barrier(pool.mem, 1)                            # every peer's chunk for this rank has landed in recv[rank]
        mine = recv[pool.rank]
        ptrs = [(flat.data_ptr() + pool.rank * n * 4) if q == pool.rank else (mine.data_ptr() + q * n * 4)
                for q in range(pool.size)]
        _launch_reduce_scatter(ptrs, 0, output_tensor, n, scale, self.max_ctas)
        return None


def install(model: torch.nn.Module, group: dist.ProcessGroup, device, mem=None, all_gather: bool = True,
            reduce_scatter: bool = True, max_ctas: int = 32, mode: str = "pull", direct: bool = True) -> _PeerPool:
    """Give every FSDP2 module group of `model` the peer-memory collectives (call after `fully_shard`).
    mode "pull": tn_peer_* kernels read the peers' buffers; mode "push": copy-engine pushes + a local reduce kernel."""
    from torch.distributed.fsdp import FSDPModule
    if mode not in ("pull", "push"):
        raise ValueError(f"fsdp_comm.install: unknown mode {mode!r}")
    pool = _PeerPool(group, device, mem)
    rs = PeerReduceScatter(pool, max_ctas) if mode == "pull" else PushReduceScatter(pool, max_ctas, direct=direct)
    ag = PeerAllGather(pool, max_ctas) if mode == "pull" else PushAllGather(pool)
    for m in model.modules():
        if isinstance(m, FSDPModule):
            if reduce_scatter:
                m.set_custom_reduce_scatter(rs)
            if all_gather:
                m.set_custom_all_gather(ag)
    return pool
