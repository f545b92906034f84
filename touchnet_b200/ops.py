"""Host-side op layer: thin, allocation-only wrappers around the C ABI (include/touchnet_b200.h) and the autograd
Functions built from them.  PyTorch is plumbing here (device memory, streams, autograd bookkeeping); every FLOP and
every byte moved on the hot path happens inside libtouchnet_b200.so.  There is no fallback: a missing library or a
non-CUDA tensor raises.

Reference call sites each op replaces are listed in include/touchnet_b200.h.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch

from . import _lib

BF16 = torch.bfloat16


def _st() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: torch.Tensor, name: str, dtype=None):
    if not t.is_cuda:
        raise _lib.TouchNetB200Error(f"{name} must be a CUDA tensor (touchnet_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise _lib.TouchNetB200Error(f"{name} must be {dtype}, got {t.dtype}")


def _rows2d(t: torch.Tensor) -> torch.Tensor:
    """View as [rows, cols] with unit inner stride (no copy unless the input is not row-contiguous)."""
    t2 = t.reshape(-1, t.shape[-1])
    if t2.stride(-1) != 1 or (t2.shape[0] > 1 and t2.stride(0) < t2.shape[1]):
        t2 = t2.contiguous()
    return t2


# ---------------------------------------------------------------------------------------------------------------
# bf16 working copies of fp32 master weights (what FSDP2's MixedPrecisionPolicy does at all-gather time,
# ref: touchnet/models/helper_func.py:163; at 1 GPU the reference stays fp32, SURVEY 9.1 - we compute in bf16 always)
# ---------------------------------------------------------------------------------------------------------------
def cast_bf16(src: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _chk(src, "src", torch.float32)
    src = src.contiguous()
    if out is None:
        out = torch.empty(src.shape, dtype=BF16, device=src.device)
    _lib.call("tn_cast_f32_bf16", src.data_ptr(), out.data_ptr(), src.numel(), _st())
    torch.autograd.graph.increment_version(out)   # raw-pointer write: a graph that saved `out` for backward must notice
    return out


def bf16_weight(w: torch.Tensor) -> torch.Tensor:
    """bf16 working copy of a parameter: the tensor itself if already bf16 (FSDP2 mixed precision hands the modules
    bf16 unsharded parameters), else a cast of the fp32 master cached ON THE PARAMETER OBJECT (attribute `_tn_bf16` =
    (version, tensor, event, epoch), keyed by the parameter's version counter) so that it lives and dies with the parameter and
    one step casts once.  Call it with the nn.Parameter itself (module level), not from inside an autograd Function.
    If the copy was produced ahead of time on the side stream (`prefetch_bf16_weights`), the current stream is made to
    wait for it here."""
    if w.dtype == BF16:
        w = w.detach()
        return w if w.is_contiguous() else w.contiguous()
    ent = getattr(w, "_tn_bf16", None)
    ok = ent is not None and ent[1].shape == w.shape and ent[1].device == w.device
    if ok and ent[0] == w._version and ent[3] == _CACHE_EPOCH:
        if ent[2] is not None:
            torch.cuda.current_stream().wait_event(ent[2])
            w._tn_bf16 = (ent[0], ent[1], None, ent[3])
        return ent[1]
    out = cast_bf16(w.detach(), ent[1] if ok else None)
    w._tn_bf16 = (w._version, out, None, _CACHE_EPOCH)
    return out


# FSDP2 refills its unsharded fp32 parameters (param_dtype=None) under `_unsafe_preserve_version_counter`, so the version
# key alone would serve stale copies there: models that find themselves FSDP-managed bump this epoch every forward.
_CACHE_EPOCH = 0


def begin_forward(module: torch.nn.Module) -> None:
    """Called at the top of the model forwards: starts a new cache epoch when parameters are FSDP-managed (casts happen
    once per forward, like the all-gather), else issues the side-stream prefetch of stale fp32 master weights."""
    global _CACHE_EPOCH
    try:
        from torch.distributed.fsdp import FSDPModule
        managed = any(isinstance(m, FSDPModule) for m in module.modules())
    except ImportError:
        managed = False
    if managed:
        _CACHE_EPOCH += 1
    else:
        prefetch_bf16_weights(module)


def invalidate_bf16_cache(module: torch.nn.Module) -> None:
    """Force a fresh fp32->bf16 cast on next use (bench.py calls this every step: the cast is part of the step, as the
    all-gather-time cast is under FSDP2).  The buffers are kept and overwritten."""
    for p in module.parameters():
        ent = getattr(p, "_tn_bf16", None)
        if ent is not None:
            p._tn_bf16 = (-1, ent[1], None, _CACHE_EPOCH)


_CAST_STREAM: dict = {}


def _is_dtensor(t) -> bool:
    try:
        from torch.distributed.tensor import DTensor
    except Exception:       # torch built without distributed
        return False
    return isinstance(t, DTensor)


def prefetch_bf16_weights(module: torch.nn.Module) -> int:
    """Issue the fp32->bf16 casts of every stale >=2-D fp32 parameter of `module` on a side stream, in parameter order,
    so that these HBM-bound copies overlap the tensor-bound GEMMs of the layers in front of them (the same idea as
    FSDP2's all-gather prefetch).  Consumers synchronise per parameter through an event (see bf16_weight).
    Returns the number of casts issued."""
    params, seen = [], set()
    for m in module.modules():          # nn.Linear weights are the only tensors consumed through bf16_weight()
        w = getattr(m, "weight", None) if isinstance(m, torch.nn.Linear) else None
        if w is None or w.dtype != torch.float32 or not w.is_cuda or id(w) in seen:
            continue
        if _is_dtensor(w) or w.untyped_storage().size() == 0:
            continue                    # FSDP2/TP-managed: the all-gather delivers bf16 (MixedPrecisionPolicy), nothing to cast
        seen.add(id(w))
        params.append(w)
    todo = []
    for p in params:
        ent = getattr(p, "_tn_bf16", None)
        ok = ent is not None and ent[1].shape == p.shape and ent[1].device == p.device
        if not (ok and ent[0] == p._version and ent[3] == _CACHE_EPOCH):
            todo.append((p, ent[1] if ok else None))
    if not todo:
        return 0
    dev = todo[0][0].device
    side = _CAST_STREAM.get(dev)
    if side is None:
        side = _CAST_STREAM[dev] = torch.cuda.Stream(device=dev)
    cur = torch.cuda.current_stream(dev)
    side.wait_stream(cur)          # everything that still reads the old copies (or writes the masters) is ahead of us
    with torch.cuda.stream(side):
        for p, buf in todo:
            out = cast_bf16(p.detach(), buf)
            out.record_stream(cur)
            ev = torch.cuda.Event()
            ev.record(side)
            p._tn_bf16 = (p._version, out, ev, _CACHE_EPOCH)
    return len(todo)


# ---------------------------------------------------------------------------------------------------------------
# raw ops (no autograd)
# ---------------------------------------------------------------------------------------------------------------
def gemm(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False, out_f32: bool = False,
         residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, M=None, N=None, K=None):
    """D[M,N] = op(A)·op(B) (+R).  a: [M,K] (a_mn=False) or [K,M] (a_mn=True); b: [N,K] or [K,N] (b_mn=True)."""
    _chk(a, "a", BF16); _chk(b, "b", BF16)
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    if M is None:
        M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
        N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
        assert K == Kb, (a.shape, b.shape, a_mn, b_mn)
    dt = torch.float32 if out_f32 else BF16
    if out is None:
        out = torch.empty((M, N), dtype=dt, device=a.device)
    assert out.dtype == dt and out.stride(1) == 1
    if residual is not None:
        assert residual.dtype == dt and residual.stride(-1) == 1
        residual = residual.reshape(M, N) if residual.dim() != 2 else residual
    _lib.call("tn_gemm_bf16", a.data_ptr(), a.stride(0), int(a_mn), b.data_ptr(), b.stride(0), int(b_mn),
              out.data_ptr(), out.stride(0), int(out_f32), None if residual is None else residual.data_ptr(),
              0 if residual is None else residual.stride(0), M, N, K, _st())
    return out


def gemm_swiglu(x: torch.Tensor, wg: torch.Tensor, wu: torch.Tensor, need_gu: bool = True):
    """G = x·Wgᵀ, U = x·Wuᵀ, H = silu(G)⊙U in one tcgen05 kernel (SwiGLU in the epilogue)."""
    _chk(x, "x", BF16); _chk(wg, "wg", BF16); _chk(wu, "wu", BF16)
    M, K = x.shape
    N = wg.shape[0]
    assert wg.shape == wu.shape == (N, K) and wg.stride(0) == wu.stride(0)
    h = torch.empty((M, N), dtype=BF16, device=x.device)
    g = torch.empty_like(h) if need_gu else None
    u = torch.empty_like(h) if need_gu else None
    _lib.call("tn_gemm_swiglu_bf16", x.data_ptr(), x.stride(0), wg.data_ptr(), wu.data_ptr(), wg.stride(0),
              None if g is None else g.data_ptr(), None if u is None else u.data_ptr(), h.data_ptr(), N, M, N, K, _st())
    return g, u, h


_FUSE_QKV = os.environ.get("TN_FUSED_QKV", "1") != "0"   # A/B switch for measurements


def qkv_fusable(M: int, nq: int, nkv: int) -> bool:
    """The three projections run as one CTA-pair GEMM when every segment is a whole number of 256-wide tiles."""
    return _FUSE_QKV and M >= 256 and nq % 256 == 0 and nkv % 256 == 0


def gemm_qkv_fwd(x, wq, wk, wv, rope=None):
    """qkv[M, nq+2nkv] = x·[Wq;Wk;Wv]ᵀ in one launch (weights stay separate tensors).  rope = (cos, sin) tables
    [M, 64]: RoPE is applied to the q and k columns in the GEMM epilogue."""
    M, K = x.shape
    nq, nkv = wq.shape[0], wk.shape[0]
    out = torch.empty((M, nq + 2 * nkv), dtype=BF16, device=x.device)
    _lib.call("tn_gemm_qkv_bf16", 0, x.data_ptr(), x.stride(0), wq.data_ptr(), wk.data_ptr(), wv.data_ptr(), wq.stride(0),
              out.data_ptr(), None, None, out.stride(0), 0, nq, nkv, nkv, M, nq + 2 * nkv, K,
              None if rope is None else rope[0].data_ptr(), None if rope is None else rope[1].data_ptr(), _st())
    return out


def gemm_qkv_dgrad(dqkv, wq, wk, wv):
    """dx[M, d] = dq·Wq + dk·Wk + dv·Wv with dq|dk|dv side by side in `dqkv`."""
    M, Kt = dqkv.shape
    d = wq.shape[1]
    out = torch.empty((M, d), dtype=BF16, device=dqkv.device)
    _lib.call("tn_gemm_qkv_bf16", 1, dqkv.data_ptr(), dqkv.stride(0), wq.data_ptr(), wk.data_ptr(), wv.data_ptr(),
              wq.stride(0), out.data_ptr(), None, None, out.stride(0), 0, wq.shape[0], wk.shape[0], wv.shape[0], M, d, Kt,
              None, None, _st())
    return out


def gemm_qkv_wgrad(dqkv, x, f32: bool, nq: int | None = None, nkv: int | None = None):
    """(dWq, dWk, dWv) = (dqᵀ·x, dkᵀ·x, dvᵀ·x) in one launch; fp32 outputs for fp32 master weights, else bf16.
    Segment sizes default to the GQA split implied by the buffer width (nq + 2*nkv) when nq is given."""
    Mred, Mt = dqkv.shape
    d = x.shape[1]
    if nq is None:
        raise _lib.TouchNetB200Error("gemm_qkv_wgrad needs the q segment width")
    nkv = (Mt - nq) // 2 if nkv is None else nkv
    dt = torch.float32 if f32 else BF16
    outs = [torch.empty((n, d), dtype=dt, device=x.device) for n in (nq, nkv, nkv)]
    _lib.call("tn_gemm_qkv_bf16", 2, dqkv.data_ptr(), dqkv.stride(0), x.data_ptr(), None, None, x.stride(0),
              outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), d, int(f32), nq, nkv, nkv, Mt, d, Mred, None,
              None, _st())
    return outs


# Measured in the 32-layer step (profiles/r02_bench_n1_dswiglu_fused.json): the fused launch runs at 828 TFLOP/s (its epilogue
# - strided G/U loads, 2 exp per element, two stores - outlasts the K=4096 mainloop it should hide behind) against 1400 for the
# plain dgrad + a 0.22 ms SwiGLU-backward pass, i.e. 0.25 ms per layer SLOWER.  Kept (bit-identical, tested) behind
# TN_FUSED_DSWIGLU=1; the default is the two-kernel path until the epilogue reads G/U through TMA.
_FUSE_DSWIGLU = os.environ.get("TN_FUSED_DSWIGLU", "0") != "0"


def gemm_dswiglu(dy: torch.Tensor, wd: torch.Tensor, g: torch.Tensor, u: torch.Tensor):
    """(dG, dU) of hm = silu(g)*u given dy = d(hm . Wd^T): the down-proj dgrad GEMM dH = dy . Wd with the SwiGLU backward
    in its epilogue (dH never written).  Falls back to the two-kernel path for shapes below one CTA-pair tile."""
    M, K = dy.shape
    N = wd.shape[1]
    if not (_FUSE_DSWIGLU and M >= 256 and N >= 256 and N % 8 == 0 and g.stride(0) == u.stride(0)):
        return swiglu_bwd(g, u, gemm(dy, wd, b_mn=True))
    dg = torch.empty_like(g)
    du = torch.empty_like(u)
    assert dg.stride(0) == du.stride(0)
    _lib.call("tn_gemm_dswiglu_bf16", dy.data_ptr(), dy.stride(0), wd.data_ptr(), wd.stride(0), g.data_ptr(), u.data_ptr(),
              g.stride(0), dg.data_ptr(), du.data_ptr(), dg.stride(0), M, N, K, _st())
    return dg, du


def swiglu_bwd(g, u, dh, dg_out=None, du_out=None):
    M, N = g.shape
    dg = torch.empty_like(g) if dg_out is None else dg_out
    du = torch.empty_like(u) if du_out is None else du_out
    _lib.call("tn_swiglu_bwd_bf16", g.data_ptr(), u.data_ptr(), dh.data_ptr(), dg.data_ptr(), du.data_ptr(), M, N,
              g.stride(0), _st())
    return dg, du


def rmsnorm_fwd(x: torch.Tensor, w: torch.Tensor, eps: float, residual: Optional[torch.Tensor] = None):
    """Returns (y, s, rstd): s = x (+ residual) is the tensor that was normalised."""
    _chk(x, "x", BF16)
    rows, d = x.shape
    y = torch.empty_like(x)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    s = torch.empty_like(x) if residual is not None else x
    w = w.detach()
    _lib.call("tn_rmsnorm_fwd_bf16", x.data_ptr(), None if residual is None else residual.data_ptr(), w.data_ptr(),
              int(w.dtype == torch.float32), s.data_ptr() if residual is not None else None, y.data_ptr(),
              rstd.data_ptr(), rows, d, float(eps), _st())
    return y, s, rstd


_norm_partials = None


def rmsnorm_bwd(s: torch.Tensor, dy: torch.Tensor, w: torch.Tensor, rstd: torch.Tensor,
                ds_extra: Optional[torch.Tensor] = None):
    """Returns (ds, dw) with ds = d/ds RMSNorm (+ ds_extra: the gradient arriving through the residual branch)."""
    global _norm_partials
    if _norm_partials is None:
        _norm_partials = _lib.load().tn_rmsnorm_bwd_num_partials()
    rows, d = s.shape
    ds = torch.empty_like(s)
    part = torch.empty((_norm_partials, d), dtype=torch.float32, device=s.device)
    dw = torch.empty(d, dtype=torch.float32, device=s.device)
    w = w.detach()
    _lib.call("tn_rmsnorm_bwd_bf16", s.data_ptr(), dy.data_ptr(), None if ds_extra is None else ds_extra.data_ptr(),
              w.data_ptr(), int(w.dtype == torch.float32), rstd.data_ptr(), ds.data_ptr(), part.data_ptr(),
              _norm_partials, dw.data_ptr(), rows, d, _st())
    return ds, dw


def rope_table(position_ids: torch.Tensor, inv_freq: torch.Tensor, scaling: float = 1.0):
    """cos/sin tables [B*T, hd/2] bf16 for per-document positions."""
    _chk(position_ids, "position_ids", torch.int64)
    pos = position_ids.reshape(-1).contiguous()
    inv = inv_freq.detach().to(device=pos.device, dtype=torch.float32).contiguous()
    rows, half = pos.numel(), inv.numel()
    cos = torch.empty((rows, half), dtype=BF16, device=pos.device)
    sin = torch.empty_like(cos)
    _lib.call("tn_rope_table", pos.data_ptr(), inv.data_ptr(), float(scaling), cos.data_ptr(), sin.data_ptr(), rows, half, _st())
    return cos, sin


def rope_apply_(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, n_heads: int, head_dim: int, inverse: bool = False):
    """In-place RoPE on x [rows, n_heads*head_dim] (row stride arbitrary)."""
    _chk(x, "x", BF16)
    rows = x.shape[0]
    assert x.stride(1) == 1 and cos.shape == (rows, head_dim // 2)
    _lib.call("tn_rope_apply_bf16", x.data_ptr(), x.stride(0), cos.data_ptr(), sin.data_ptr(), rows, n_heads, head_dim,
              int(inverse), _st())
    return x


class AttnPlan:
    """Per-step packing metadata shared by every layer: int32 document ids + per-block kv/q ranges, built on device."""

    def __init__(self, doc_ids: torch.Tensor, Tq: Optional[int] = None, q_blk_off: int = 0):
        """doc_ids: GLOBAL [B,T] document ids.  Tq / q_blk_off: context-parallel query window (rows
        [q_blk_off*128, q_blk_off*128 + Tq) of the sequence live on this rank); default = the whole sequence."""
        _chk(doc_ids, "attention_mask (document ids)")
        assert doc_ids.dim() == 2
        self.B, self.T = doc_ids.shape
        self.Tq = self.T if Tq is None else int(Tq)
        self.q_blk_off = int(q_blk_off)
        self.cp_group = None
        self.tp = None            # tensor_parallel.TPContext when the block runs tensor + sequence parallel
        self.doc = doc_ids.to(torch.int32).contiguous()
        n_ints = int(_lib.load().tn_attn_meta_ints(self.B, self.T))
        self.meta = torch.empty(n_ints, dtype=torch.int32, device=doc_ids.device)
        _lib.call("tn_attn_prep", self.doc.data_ptr(), self.meta.data_ptr(), self.B, self.T, _st())


def attn_fwd(q, k, v, plan: AttnPlan, H: int, KV: int, scale: float):
    """q [B*Tq, H*128], k/v [B*T, KV*128] (row strides arbitrary) -> o [B*Tq, H*128], lse [B,H,Tq].
    Tq == T unless the plan describes a context-parallel query window (plan.Tq, plan.q_blk_off)."""
    B, T = plan.B, plan.T
    Tq = plan.Tq
    o = torch.empty((B * Tq, H * 128), dtype=BF16, device=q.device)
    lse = torch.empty((B, H, Tq), dtype=torch.float32, device=q.device)
    _lib.call("tn_attn_fwd_bf16", q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
              o.data_ptr(), o.stride(0), lse.data_ptr(), plan.doc.data_ptr(), plan.meta.data_ptr(), B, T, H, KV,
              float(scale), Tq, plan.q_blk_off, _st())
    return o, lse


def attn_bwd(q, k, v, o, do, lse, plan: AttnPlan, H: int, KV: int, scale: float, out=None, rope=None):
    """`out` = (dq, dk, dv) preallocated (possibly strided views of one [B*T, (H+2KV)*128] buffer).
    `rope` = (cos, sin) tables: the kernels' epilogues apply the inverse rotation, i.e. dq/dk come back w.r.t. the
    un-rotated projections (only without context parallelism, where dK rows and table rows coincide)."""
    B, T = plan.B, plan.T
    Tq = plan.Tq
    if out is None:
        dq = torch.empty((B * Tq, H * 128), dtype=BF16, device=q.device)
        dk = torch.empty((B * T, KV * 128), dtype=BF16, device=q.device)
        dv = torch.empty_like(dk)
    else:
        dq, dk, dv = out
    delta = torch.empty((B, H, Tq), dtype=torch.float32, device=q.device)
    if do.stride(1) != 1:
        do = do.contiguous()
    _lib.call("tn_attn_bwd_bf16", q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
              o.data_ptr(), o.stride(0), do.data_ptr(), do.stride(0), lse.data_ptr(), delta.data_ptr(),
              dq.data_ptr(), dq.stride(0), dk.data_ptr(), dk.stride(0), dv.data_ptr(), dv.stride(0),
              plan.doc.data_ptr(), plan.meta.data_ptr(), B, T, H, KV, float(scale), Tq, plan.q_blk_off,
              None if rope is None else rope[0].data_ptr(), None if rope is None else rope[1].data_ptr(),
              None if rope is None else rope[0].data_ptr(), None if rope is None else rope[1].data_ptr(), _st())
    return dq, dk, dv


def embed_add(input_ids: Optional[torch.Tensor], embed: Optional[torch.Tensor], proj: Optional[torch.Tensor],
              rows: int, d: int, nan_flag: Optional[torch.Tensor] = None):
    """E = embed[input_ids] + proj (either may be None)."""
    dev = (proj if proj is not None else input_ids).device
    e = torch.empty((rows, d), dtype=BF16, device=dev)
    emb = None if embed is None else embed.detach()
    _lib.call("tn_embed_add_bf16", None if input_ids is None else input_ids.data_ptr(),
              None if emb is None else emb.data_ptr(), int(emb is not None and emb.dtype == torch.float32),
              None if proj is None else proj.data_ptr(), e.data_ptr(),
              None if nan_flag is None else nan_flag.data_ptr(), rows, d, 0 if emb is None else emb.shape[0], _st())
    return e


# ---------------------------------------------------------------------------------------------------------------
# autograd Functions
# ---------------------------------------------------------------------------------------------------------------
def _wgrad(dy2: torch.Tensor, x2: torch.Tensor, f32: bool) -> torch.Tensor:
    """dW[N,K] = dyᵀ·x, in the parameter's dtype (fp32 master -> fp32 gradient straight from the TMEM accumulator)."""
    return gemm(dy2, x2, a_mn=True, b_mn=True, out_f32=f32)


class LinearFn(torch.autograd.Function):
    """y = x·Wᵀ (+ bias) (+ residual).  F.linear at hf:modeling_llama.py:251-289,182-184;
    ref: touchnet/models/touch_audio/modeling_touch_audio.py:127 (projector).
    `w` is the parameter (gradient target, master dtype), `wb` its bf16 working copy (non-differentiable)."""

    @staticmethod
    def forward(ctx, x, w, wb, bias, residual):
        x2 = _rows2d(x)
        if x2.dtype != BF16:
            x2 = x2.to(BF16)
        r2 = None if residual is None else _rows2d(residual)
        y = gemm(x2, wb, residual=r2)
        if bias is not None:
            y += bias.to(BF16)
        ctx.save_for_backward(x2, wb)
        ctx.w_dtype = w.dtype
        ctx.has_bias = bias is not None
        ctx.has_res = residual is not None
        ctx.x_shape = x.shape
        ctx.x_dtype = x.dtype
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, wb = ctx.saved_tensors
        dy2 = _rows2d(dy)
        if dy2.dtype != BF16:
            dy2 = dy2.to(BF16)
        dx = dw = db = dr = None
        if ctx.needs_input_grad[0]:
            dx = gemm(dy2, wb, b_mn=True).view(ctx.x_shape).to(ctx.x_dtype)
        if ctx.needs_input_grad[1]:
            dw = gemm(dy2, x2, a_mn=True, b_mn=True, out_f32=(ctx.w_dtype == torch.float32))
        if ctx.has_bias and ctx.needs_input_grad[3]:
            db = dy2.float().sum(0).to(ctx.w_dtype)
        if ctx.has_res and ctx.needs_input_grad[4]:
            dr = dy
        return dx, dw, None, db, dr


def linear(x, w, bias=None, residual=None):
    return LinearFn.apply(x, w, bf16_weight(w), bias, residual)


class RMSNormFn(torch.autograd.Function):
    """hf: LlamaRMSNorm.forward modeling_llama.py:62-67."""

    @staticmethod
    def forward(ctx, x, w, eps):
        x2 = _rows2d(x)
        y, _, rstd = rmsnorm_fwd(x2, w, eps)
        ctx.save_for_backward(x2, w, rstd)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, w, rstd = ctx.saved_tensors
        ds, dw = rmsnorm_bwd(x2, _rows2d(dy), w, rstd)
        return ds.view(dy.shape), dw.to(w.dtype), None


def rms_norm(x, w, eps):
    return RMSNormFn.apply(x, w, eps)


class PackedAttentionFn(torch.autograd.Function):
    """RoPE + block-causal document attention on projected q/k/v ([B*T, heads*128] GEMM outputs, rotated in place).
    hf: apply_rotary_pos_emb :151-168 + flex_attention_forward (integrations/flex_attention.py:262-364).
    q and k are OVERWRITTEN with their rotated values: hand in tensors nothing else reads (the projection outputs; an
    nn.Linear does not need its own output for backward).  The decoder block does not use this node (it fuses RoPE into
    the QKV GEMM epilogue); it is the stand-alone form for callers that keep separate projections."""

    @staticmethod
    def forward(ctx, q, k, v, cos, sin, plan, H, KV, scale):
        rope_apply_(q, cos, sin, H, 128)
        rope_apply_(k, cos, sin, KV, 128)
        o, lse = attn_fwd(q, k, v, plan, H, KV, scale)
        ctx.save_for_backward(q, k, v, o, lse, cos, sin)
        ctx.plan, ctx.H, ctx.KV, ctx.scale = plan, H, KV, scale
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse, cos, sin = ctx.saved_tensors
        dq, dk, dv = attn_bwd(q, k, v, o, do, lse, ctx.plan, ctx.H, ctx.KV, ctx.scale, rope=(cos, sin))
        return dq, dk, dv, None, None, None, None, None, None


class AttentionFn(torch.autograd.Function):
    """Document-masked attention on already-rotated q/k (the HF attention-interface seam hands RoPE'd tensors)."""

    @staticmethod
    def forward(ctx, q, k, v, plan, H, KV, scale):
        o, lse = attn_fwd(q, k, v, plan, H, KV, scale)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.plan, ctx.H, ctx.KV, ctx.scale = plan, H, KV, scale
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse = ctx.saved_tensors
        dq, dk, dv = attn_bwd(q, k, v, o, do, lse, ctx.plan, ctx.H, ctx.KV, ctx.scale)
        return dq, dk, dv, None, None, None, None


class SwiGLUFn(torch.autograd.Function):
    """hmid = silu(x·Wgᵀ) ⊙ (x·Wuᵀ), one kernel.  hf: LlamaMLP.forward modeling_llama.py:182-184."""

    @staticmethod
    def forward(ctx, x, wg, wu, wgb, wub):
        x2 = _rows2d(x)
        g, u, h = gemm_swiglu(x2, wgb, wub)
        ctx.save_for_backward(x2, wgb, wub, g, u)
        ctx.f32 = wg.dtype == torch.float32
        return h.view(*x.shape[:-1], wg.shape[0])

    @staticmethod
    def backward(ctx, dh):
        x2, wgb, wub, g, u = ctx.saved_tensors
        dg, du = swiglu_bwd(g, u, _rows2d(dh))
        dx = gemm(dg, wgb, b_mn=True)
        dx = gemm(du, wub, b_mn=True, residual=dx, out=dx)
        return dx.view(*dh.shape[:-1], x2.shape[1]), _wgrad(dg, x2, ctx.f32), _wgrad(du, x2, ctx.f32), None, None


def swiglu_mlp_in(x, wg, wu):
    return SwiGLUFn.apply(x, wg, wu, bf16_weight(wg), bf16_weight(wu))


def _pad_heads(x: torch.Tensor, heads: int, hd: int) -> torch.Tensor:
    """[rows, heads*hd] -> [rows, heads*128] with every head zero-padded to the 128 columns the attention kernels are built
    for: zero q/k columns add nothing to the scores, zero v columns give zero output columns - the result is exact."""
    out = torch.zeros((x.shape[0], heads, 128), dtype=x.dtype, device=x.device)
    out[:, :, :hd] = x.reshape(x.shape[0], heads, hd)
    return out.view(x.shape[0], heads * 128)


def _unpad_heads(x: torch.Tensor, heads: int, hd: int) -> torch.Tensor:
    return x.view(x.shape[0], heads, 128)[:, :, :hd].reshape(x.shape[0], heads * hd)


class DecoderLayerFn(torch.autograd.Function):
    """One pre-norm decoder block as a single autograd node (hf: LlamaDecoderLayer.forward modeling_llama.py:303-333):
    7 kernel launches forward, residual adds fused into the o_proj / down_proj GEMM epilogues, q/k/v projections one
    segmented GEMM, the residual-branch gradient fused into the RMSNorm backward kernel.  Parameters stay separate
    nn.Linear weights (HF FQNs) so FSDP2 / DCP / convert_* of the reference keep working
    (ref: touchnet/models/helper_func.py:134-202).  Use `decoder_layer(...)`: it resolves the bf16 working copies of the
    fp32 master weights on the parameter objects and passes them in as non-differentiable inputs."""

    @staticmethod
    def forward(ctx, x, ln1, wq, wk, wv, bq, bk, bv, wo, ln2, wg, wu, wd, wqb, wkb, wvb, wob, wgb, wub, wdb, cos, sin,
                plan, H, KV, eps, hd=128):
        x2 = _rows2d(x)
        scale = 1.0 / math.sqrt(hd)
        if hd != 128 and (hd > 128 or hd % 8 or plan.tp is not None or plan.cp_group is not None):
            raise _lib.TouchNetB200Error(f"head_dim {hd}: only multiples of 8 up to 128, without tensor / context parallelism")
        tp = plan.tp                # tensor + sequence parallel: x holds T/tp rows, weights are this rank's shards
        if tp is not None and plan.cp_group is not None:
            raise _lib.TouchNetB200Error("tensor parallelism and context parallelism cannot be combined in one block")
        h1, _, rstd1 = rmsnorm_fwd(x2, ln1, eps)
        if tp is not None:
            h1 = tp.gather_rows(h1, keep=True)                               # [B*T, d]: every row, for the local heads
        nq, nkv = wq.shape[0], wk.shape[0]
        rope_in_gemm = qkv_fusable(h1.shape[0], nq, nkv) and bq is None and hd == 128   # RoPE in the QKV GEMM epilogue
        if qkv_fusable(h1.shape[0], nq, nkv):
            qkv = gemm_qkv_fwd(h1, wqb, wkb, wvb, rope=(cos, sin) if rope_in_gemm else None)
            q, k, v = qkv[:, :nq], qkv[:, nq:nq + nkv], qkv[:, nq + nkv:]
        else:
            q = gemm(h1, wqb); k = gemm(h1, wkb); v = gemm(h1, wvb)
        if bq is not None:
            q += bq.to(BF16); k += bk.to(BF16); v += bv.to(BF16)
        if not rope_in_gemm:
            rope_apply_(q, cos, sin, H, hd)
            rope_apply_(k, cos, sin, KV, hd)
        if hd != 128:      # head_dim 64 (Llama-3.2-1B, the reference's example config) etc.: zero-padded heads, 128-wide kernels
            q, k, v = _pad_heads(q, H, hd), _pad_heads(k, KV, hd), _pad_heads(v, KV, hd)
        if plan.cp_group is None:
            o, lse = attn_fwd(q, k, v, plan, H, KV, scale)
        else:   # context parallel: K/V of all cp ranks are gathered, the kernel runs on this rank's query window
            from . import context_parallel as _cp
            o, lse, k, v = _cp.cp_attn_fwd(q, k, v, plan, H, KV, scale)
        if tp is None:
            x1 = gemm(o if hd == 128 else _unpad_heads(o, H, hd), wob, residual=x2)
        else:                                                                # partial sums over tp -> this rank's rows
            x1 = tp.gemm_reduce_scatter(o, wob, x2)
        h2, _, rstd2 = rmsnorm_fwd(x1, ln2, eps)
        if tp is not None:
            h2 = tp.gather_rows(h2, keep=True)
        g, u, hm = gemm_swiglu(h2, wgb, wub)
        if tp is None:
            out = gemm(hm, wdb, residual=x1)
        else:
            out = tp.gemm_reduce_scatter(hm, wdb, x1)
        ctx.save_for_backward(x2, ln1, ln2, wqb, wkb, wvb, wob, wgb, wub, wdb, cos, sin, rstd1, h1, q, k, v, o, lse, x1,
                              rstd2, h2, g, u, hm)
        ctx.plan, ctx.H, ctx.KV, ctx.scale, ctx.has_bias, ctx.hd = plan, H, KV, scale, bq is not None, hd
        ctx.f32 = wq.dtype == torch.float32
        ctx.w_dtype = wq.dtype
        return out.view(x.shape)

    @staticmethod
    def backward(ctx, dout):
        (x2, ln1, ln2, wqb, wkb, wvb, wob, wgb, wub, wdb, cos, sin, rstd1, h1, q, k, v, o, lse, x1, rstd2, h2, g, u,
         hm) = ctx.saved_tensors
        H, KV, f32 = ctx.H, ctx.KV, ctx.f32
        tp = ctx.plan.tp
        d2 = _rows2d(dout)
        if d2.dtype != BF16:
            d2 = d2.to(BF16)
        # ---- MLP ----
        d2f = d2 if tp is None else tp.gather_rows(d2)     # backward of the forward reduce-scatter
        dg, du = gemm_dswiglu(d2f, wdb, g, u)          # down-proj dgrad, SwiGLU backward in its epilogue
        dwd = _wgrad(d2f, hm, f32)
        del d2f
        dh2 = gemm(dg, wgb, b_mn=True)
        dh2 = gemm(du, wub, b_mn=True, residual=dh2, out=dh2)
        dwg = _wgrad(dg, h2, f32)
        dwu = _wgrad(du, h2, f32)
        del dg, du
        if tp is not None:
            dh2 = tp.reduce_scatter(dh2)               # backward of the forward all-gather
        dx1, dln2 = rmsnorm_bwd(x1, dh2, ln2, rstd2, ds_extra=d2)
        # ---- attention ----
        dx1f = dx1 if tp is None else tp.gather_rows(dx1)
        hd = ctx.hd
        do = gemm(dx1f, wob, b_mn=True)
        dwo = _wgrad(dx1f, o if hd == 128 else _unpad_heads(o, H, hd), f32)
        if hd != 128:
            do = _pad_heads(do, H, hd)
        del dx1f
        nq, nkv = wqb.shape[0], wkb.shape[0]
        fused = qkv_fusable(h1.shape[0], nq, nkv)
        if ctx.plan.cp_group is not None:
            from . import context_parallel as _cp
            dq, dk, dv = _cp.cp_attn_bwd(q, k, v, o, do, lse, ctx.plan, H, KV, ctx.scale)   # dK/dV reduce-scattered
            if fused:
                dqkv = torch.cat([dq, dk, dv], dim=1)
                dq, dk, dv = dqkv[:, :nq], dqkv[:, nq:nq + nkv], dqkv[:, nq + nkv:]
            rope_apply_(dq, cos, sin, H, 128, inverse=True)
            rope_apply_(dk, cos, sin, KV, 128, inverse=True)
        elif hd != 128:   # padded heads: gradients of the padded tensors, un-padded, then the inverse rotation at head_dim hd
            dqp, dkp, dvp = attn_bwd(q, k, v, o, do, lse, ctx.plan, H, KV, ctx.scale)
            dq, dk, dv = _unpad_heads(dqp, H, hd), _unpad_heads(dkp, KV, hd), _unpad_heads(dvp, KV, hd)
            rope_apply_(dq, cos, sin, H, hd, inverse=True)
            rope_apply_(dk, cos, sin, KV, hd, inverse=True)
            if fused:
                dqkv = torch.cat([dq, dk, dv], dim=1)
                dq, dk, dv = dqkv[:, :nq], dqkv[:, nq:nq + nkv], dqkv[:, nq + nkv:]
        elif fused:   # inverse RoPE happens in the attention-backward epilogues
            dqkv = torch.empty((h1.shape[0], nq + 2 * nkv), dtype=BF16, device=x2.device)
            dq, dk, dv = dqkv[:, :nq], dqkv[:, nq:nq + nkv], dqkv[:, nq + nkv:]
            attn_bwd(q, k, v, o, do, lse, ctx.plan, H, KV, ctx.scale, out=(dq, dk, dv), rope=(cos, sin))
        else:
            dq, dk, dv = attn_bwd(q, k, v, o, do, lse, ctx.plan, H, KV, ctx.scale, rope=(cos, sin))
        if fused:
            dh1 = gemm_qkv_dgrad(dqkv, wqb, wkb, wvb)
            dwq, dwk, dwv = gemm_qkv_wgrad(dqkv, h1, f32, nq, nkv)
        else:
            dh1 = gemm(dq, wqb, b_mn=True)
            dh1 = gemm(dk, wkb, b_mn=True, residual=dh1, out=dh1)
            dh1 = gemm(dv, wvb, b_mn=True, residual=dh1, out=dh1)
            dwq = _wgrad(dq, h1, f32); dwk = _wgrad(dk, h1, f32); dwv = _wgrad(dv, h1, f32)
        dbq = dbk = dbv = None
        if ctx.has_bias:
            dbq = dq.float().sum(0).to(ctx.w_dtype); dbk = dk.float().sum(0).to(ctx.w_dtype); dbv = dv.float().sum(0).to(ctx.w_dtype)
        if tp is not None:
            dh1 = tp.reduce_scatter(dh1)
        dx, dln1 = rmsnorm_bwd(x2, dh1, ln1, rstd1, ds_extra=dx1)
        if tp is not None:                             # replicated norm weights, sequence-sharded rows: sum the partials
            tp.all_reduce_(dln1)
            tp.all_reduce_(dln2)
        return (dx.view(dout.shape), dln1.to(ln1.dtype), dwq, dwk, dwv, dbq, dbk, dbv, dwo, dln2.to(ln2.dtype), dwg, dwu,
                dwd) + (None,) * 14


def decoder_layer(x, ln1, wq, wk, wv, bq, bk, bv, wo, ln2, wg, wu, wd, cos, sin, plan, H, KV, eps, hd=128):
    return DecoderLayerFn.apply(x, ln1, wq, wk, wv, bq, bk, bv, wo, ln2, wg, wu, wd, bf16_weight(wq), bf16_weight(wk),
                                bf16_weight(wv), bf16_weight(wo), bf16_weight(wg), bf16_weight(wu), bf16_weight(wd),
                                cos, sin, plan, H, KV, eps, hd)
