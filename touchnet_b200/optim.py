"""Optimizer step next to the path: AdamW and gradient-norm clipping on the rank-local fp32 shards.

    B200AdamW         drop-in for what ref: touchnet/utils/optimizer.py:127-172 builds
                      (torch.optim.AdamW(lr, betas=(0.9, 0.95), eps, weight_decay=0.1, fused=True)); same param-group
                      options, same state keys ("step", "exp_avg", "exp_avg_sq") so DCP optimizer checkpoints carry over.
    clip_grad_norm_   same contract as ref: touchnet/utils/distributed.py:426-491 (total 2-norm over all parameters incl.
                      DTensor shards, returned as a tensor; gradients scaled by max_norm / (norm + 1e-6) when above).

Fused form (what a B200 step should do): `clip_grad_norm_(params, max_norm, defer_to=optimizer)` only computes the norm
and hands the clip coefficient to the optimizer as a DEVICE scalar; `optimizer.step()` then applies it while it reads the
gradient, and writes the bf16 working copy of every updated nn.Linear weight in the same pass (ops.bf16_weight finds it
fresh, so the next forward has no cast at all).  No host synchronisation anywhere.
"""
from __future__ import annotations

import ctypes
import math
from typing import Iterable, Optional, Union

import torch
import torch.distributed as dist

from . import _lib, ops

_n_partials = None


def _local(t: torch.Tensor) -> torch.Tensor:
    return t.to_local() if ops._is_dtensor(t) else t


def _sumsq(x: torch.Tensor) -> torch.Tensor:
    """Sum of squares of a local fp32 tensor as a 0-d device tensor (deterministic two-stage reduction)."""
    global _n_partials
    if _n_partials is None:
        _n_partials = _lib.load().tn_sumsq_num_partials()
    ops._chk(x, "grad", torch.float32)
    if x.numel() == 0:          # an empty local shard (uneven FSDP2 sharding): contributes 0, every rank stays in step
        return torch.zeros((), dtype=torch.float32, device=x.device)
    if not x.is_contiguous() or x.data_ptr() % 16:
        x = x.contiguous().clone()   # unaligned FSDP2 gradient view: the kernel wants 16-byte vectors (read-only use)
    part = torch.empty(_n_partials, dtype=torch.float32, device=x.device)
    used = ctypes.c_int(0)
    _lib.call("tn_sumsq_f32", x.data_ptr(), x.numel(), part.data_ptr(), ctypes.byref(used), ops._st())
    return part[: used.value].sum()


def _shard_groups(t) -> list:
    """Process groups over which the local sums of squares of a DTensor must be added (its sharded mesh dims)."""
    if not ops._is_dtensor(t):
        return []
    from torch.distributed.tensor import Shard
    mesh = t.device_mesh
    return [mesh.get_group(i) for i, pl in enumerate(t.placements) if isinstance(pl, Shard)]


@torch.no_grad()
def clip_grad_norm_(parameters: Union[torch.Tensor, Iterable[torch.Tensor]], max_norm: float, norm_type: float = 2.0,
                    error_if_nonfinite: bool = False, foreach: Optional[bool] = None, pp_mesh=None,
                    defer_to: Optional["B200AdamW"] = None) -> torch.Tensor:
    """ref: touchnet/utils/distributed.py:426-491.  2-norm only (every reference config).  Returns the total norm (0-d
    device tensor, not synchronised).  With `defer_to=optimizer` the gradients are left untouched and the optimizer
    applies the clip coefficient inside its step."""
    if norm_type != 2.0:
        raise _lib.TouchNetB200Error("clip_grad_norm_: only the 2-norm is implemented (the reference's setting)")
    if pp_mesh is not None:
        raise _lib.TouchNetB200Error("clip_grad_norm_: pipeline parallelism is out of scope")
    params = [parameters] if isinstance(parameters, torch.Tensor) else list(parameters)
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return torch.zeros((), dtype=torch.float32)
    # bucket the local sums by the set of groups they still have to be reduced over (FSDP shards, TP shards, ...)
    buckets: dict = {}
    for g in grads:
        key = tuple(_shard_groups(g))
        s = _sumsq(_local(g))
        buckets[key] = s if key not in buckets else buckets[key] + s
    total = None
    for groups, s in buckets.items():
        for grp in groups:
            dist.all_reduce(s, op=dist.ReduceOp.SUM, group=grp)
        total = s if total is None else total + s
    total_norm = total.sqrt()
    if error_if_nonfinite and not bool(torch.isfinite(total_norm)):
        raise RuntimeError("The total norm for gradients is non-finite, so it cannot be clipped.")
    coef = (max_norm / (total_norm + 1e-6)).clamp(max=1.0).to(torch.float32).reshape(1)   # clip_grads_with_norm_
    if defer_to is not None:
        defer_to.grad_scale = coef
    else:
        for g in grads:
            gl = _local(g)
            if gl.numel() == 0:
                continue
            if gl.is_contiguous() and gl.data_ptr() % 16 == 0:
                _lib.call("tn_scale_f32", gl.data_ptr(), gl.numel(), coef.data_ptr(), ops._st())
                torch.autograd.graph.increment_version(gl)      # written through the raw pointer
            else:                    # unaligned / strided shard view: same arithmetic, plain in-place multiply
                gl.mul_(coef)
    return total_norm


class B200AdamW(torch.optim.Optimizer):
    """AdamW with decoupled weight decay on fp32 parameters (plain tensors or DTensor shards), one kernel per tensor."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.95), eps: float = 1e-8, weight_decay: float = 0.1,
                 write_bf16: bool = True, **ignored):
        # `fused` / `foreach` of the reference's optimizer_kwargs are accepted and ignored: this is the fused form
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.write_bf16 = write_bf16
        self.grad_scale: Optional[torch.Tensor] = None      # device scalar set by clip_grad_norm_(defer_to=self)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        gs = self.grad_scale
        for group in self.param_groups:
            lr, (b1, b2), eps, wd = float(group["lr"]), group["betas"], group["eps"], group["weight_decay"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)      # host-side counter, as torch's non-capturable path
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                t = float(st["step"])
                pl, gl = _local(p), _local(p.grad)
                if pl.numel() == 0:
                    continue
                ml, vl = _local(st["exp_avg"]), _local(st["exp_avg_sq"])
                if pl.dtype != torch.float32 or gl.dtype != torch.float32:
                    raise _lib.TouchNetB200Error("B200AdamW needs fp32 parameters and gradients (fp32 master weights)")
                if not (pl.is_contiguous() and gl.is_contiguous()):
                    raise _lib.TouchNetB200Error("B200AdamW needs contiguous parameter / gradient shards")
                pb = None
                if self.write_bf16 and pl.dim() >= 2 and not ops._is_dtensor(p):
                    ent = getattr(p, "_tn_bf16", None)
                    if ent is not None and ent[1].shape == pl.shape and ent[1].device == pl.device:
                        pb = ent[1]
                _lib.call("tn_adamw_f32", pl.data_ptr(), gl.data_ptr(), ml.data_ptr(), vl.data_ptr(),
                          None if pb is None else pb.data_ptr(), pl.numel(), lr, float(b1), float(b2), float(eps),
                          float(wd), 1.0 - b1 ** t, math.sqrt(1.0 - b2 ** t), None if gs is None else gs.data_ptr(),
                          ops._st())
                torch.autograd.graph.increment_version(pl)      # the kernel wrote through the raw pointer
                if pb is not None:
                    torch.autograd.graph.increment_version(pb)
                    p._tn_bf16 = (p._version, pb, None, ops._CACHE_EPOCH)   # fresh working copy: next forward casts nothing
        self.grad_scale = None
        return loss
