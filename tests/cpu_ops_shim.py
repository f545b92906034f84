"""TEST INFRASTRUCTURE ONLY: plain-torch stand-ins for the raw CUDA ops of touchnet_b200.ops, so that the HOST logic around
them (autograd wiring of the decoder block, tensor-/context-parallel collectives, DTensor parameter plumbing) can be
exercised on CPU with gloo at world_size 2.  Nothing in the product imports this; the product has no CPU path
(ops._chk raises).  Same call signatures and storage dtypes (bf16 activations, fp32 statistics) as the kernels."""
import math

import torch

from touchnet_b200 import ops

BF16 = torch.bfloat16


def _mm(a, b, a_mn, b_mn):
    A = a.float().t() if a_mn else a.float()
    Bm = b.float() if b_mn else b.float().t()
    return A @ Bm


def gemm(a, b, *, a_mn=False, b_mn=False, out_f32=False, residual=None, out=None, M=None, N=None, K=None):
    y = _mm(a, b, a_mn, b_mn)
    if residual is not None:
        y = y + residual.float().reshape(y.shape)
    y = y if out_f32 else y.to(BF16)
    if out is not None:
        out.copy_(y)
        return out
    return y


def gemm_swiglu(x, wg, wu, need_gu=True):
    g = (x.float() @ wg.float().t())
    u = (x.float() @ wu.float().t())
    h = (torch.nn.functional.silu(g) * u).to(BF16)
    return (g.to(BF16), u.to(BF16), h) if need_gu else (None, None, h)


def swiglu_bwd(g, u, dh, dg_out=None, du_out=None):
    gf, uf, d = g.float(), u.float(), dh.float()
    sig = torch.sigmoid(gf)
    dg = d * uf * sig * (1 + gf * (1 - sig))
    du = d * gf * sig
    return dg.to(BF16), du.to(BF16)


def rmsnorm_fwd(x, w, eps, residual=None):
    s = x if residual is None else (x.float() + residual.float()).to(BF16)
    sf = s.float()
    rstd = torch.rsqrt(sf.pow(2).mean(-1) + eps)
    y = (w.detach().float() * (sf * rstd[:, None]).to(BF16).float()).to(BF16)
    return y, s, rstd


def rmsnorm_bwd(s, dy, w, rstd, ds_extra=None):
    sf, dyf, wf = s.float(), dy.float(), w.detach().float()
    xhat = sf * rstd[:, None]
    dxhat = dyf * wf
    ds = rstd[:, None] * (dxhat - xhat * (dxhat * xhat).mean(-1, keepdim=True))
    if ds_extra is not None:
        ds = ds + ds_extra.float()
    return ds.to(BF16), (dyf * xhat).sum(0)


def rope_table(position_ids, inv_freq, scaling=1.0):
    ang = position_ids.reshape(-1).float()[:, None] * inv_freq.detach().float()[None]
    return (ang.cos() * scaling).to(BF16), (ang.sin() * scaling).to(BF16)


def rope_apply_(x, cos, sin, n_heads, head_dim, inverse=False):
    rows = x.shape[0]
    xf = x.float().reshape(rows, n_heads, head_dim)
    c, s = cos.float()[:, None, :], sin.float()[:, None, :]
    if inverse:
        s = -s
    a, b = xf[..., : head_dim // 2], xf[..., head_dim // 2:]
    y = torch.cat([a * c - b * s, b * c + a * s], dim=-1).reshape(rows, n_heads * head_dim)
    x.copy_(y.to(BF16))
    return x


def cast_bf16(src, out=None):
    y = src.detach().to(BF16)
    if out is not None:
        out.copy_(y)
        return out
    return y


def embed_add(input_ids, embed, proj, rows, d, nan_flag=None):
    e = torch.zeros((rows, d), dtype=torch.float32)
    if embed is not None:
        e = e + embed.detach().float()[input_ids]
    if proj is not None:
        e = e + proj.float()
    return e.to(BF16)


class AttnPlan:
    def __init__(self, doc_ids, Tq=None, q_blk_off=0):
        self.B, self.T = doc_ids.shape
        self.Tq = self.T if Tq is None else int(Tq)
        self.q_blk_off = int(q_blk_off)
        self.cp_group = None
        self.tp = None
        self.doc = doc_ids.to(torch.int32).contiguous()
        self.meta = _block_meta(self.doc)


def _block_meta(doc):
    """{kv_lo, kv_end, q_end, canonical} per 128-row block with the semantics of tn_attn_prep for canonical ids
    (non-decreasing runs): kv_lo = block of the start of the document of the block's first valid row."""
    B, T = doc.shape
    nblk = (T + 127) // 128
    meta = torch.zeros(B, nblk, 4, dtype=torch.int32)
    for b in range(B):
        d = doc[b].tolist()
        for blk in range(nblk):
            rows = [t for t in range(blk * 128, min(T, blk * 128 + 128)) if d[t] > 0]
            if not rows:
                continue
            t = rows[0]
            while t > 0 and d[t - 1] == d[rows[0]]:
                t -= 1
            meta[b, blk] = torch.tensor([t // 128, blk + 1, nblk, 1], dtype=torch.int32)
    return meta.reshape(-1)


def _mask(plan):
    off = plan.q_blk_off * 128
    doc = plan.doc
    qi = torch.arange(off, off + plan.Tq)
    kj = torch.arange(plan.T)
    dq = doc[:, off:off + plan.Tq]
    return (qi[None, :, None] >= kj[None, None, :]) & (dq[:, :, None] == doc[:, None, :]) & (dq[:, :, None] > 0)  # [B,Tq,T]


def _heads(x, B, rows, n):
    return x.float().reshape(B, rows, n, 128).permute(0, 2, 1, 3)


def _finite(x):
    """Rows the kernels never load may be uninitialised in the callers' buffers (context-parallel halo exchange); the dense
    stand-in multiplies them by exact zeros, so make them finite first."""
    return torch.nan_to_num(x.float(), nan=0.0, posinf=0.0, neginf=0.0)


def attn_fwd(q, k, v, plan, H, KV, scale):
    B, T, Tq = plan.B, plan.T, plan.Tq
    k, v = _finite(k), _finite(v)
    G = H // KV
    qh, kh, vh = _heads(q, B, Tq, H), _heads(k, B, T, KV).repeat_interleave(G, 1), _heads(v, B, T, KV).repeat_interleave(G, 1)
    s = (qh @ kh.transpose(-1, -2)) * scale
    m = _mask(plan)[:, None]
    s = s.masked_fill(~m, float("-inf"))
    lse = torch.logsumexp(s, dim=-1)
    dead = ~m.any(-1).expand_as(lse)
    lse = torch.where(dead, torch.full_like(lse, float("inf")), lse)
    p = torch.exp(s - lse[..., None]).masked_fill(~m, 0.0)
    o = (p @ vh).permute(0, 2, 1, 3).reshape(B * Tq, H * 128).to(BF16)
    return o, lse


def attn_bwd(q, k, v, o, do, lse, plan, H, KV, scale, out=None, rope=None):
    B, T, Tq = plan.B, plan.T, plan.Tq
    k, v = _finite(k), _finite(v)
    G = H // KV
    qh, kh, vh = _heads(q, B, Tq, H), _heads(k, B, T, KV).repeat_interleave(G, 1), _heads(v, B, T, KV).repeat_interleave(G, 1)
    oh, doh = _heads(o, B, Tq, H), _heads(do, B, Tq, H)
    m = _mask(plan)[:, None]
    s = (qh @ kh.transpose(-1, -2)) * scale
    p = torch.exp(s - lse[..., None]).masked_fill(~m, 0.0)
    p = torch.nan_to_num(p, nan=0.0)
    delta = (oh * doh).sum(-1, keepdim=True)
    dvh = p.transpose(-1, -2) @ doh
    ds = p * (doh @ vh.transpose(-1, -2) - delta)
    dqh = (ds @ kh) * scale
    dkh = (ds.transpose(-1, -2) @ qh) * scale
    dkh = dkh.reshape(B, KV, G, T, 128).sum(2)
    dvh = dvh.reshape(B, KV, G, T, 128).sum(2)
    dq = dqh.permute(0, 2, 1, 3).reshape(B * Tq, H * 128).to(BF16)
    dk = dkh.permute(0, 2, 1, 3).reshape(B * T, KV * 128).to(BF16)
    dv = dvh.permute(0, 2, 1, 3).reshape(B * T, KV * 128).to(BF16)
    if rope is not None:
        rope_apply_(dq, rope[0], rope[1], H, 128, inverse=True)
        rope_apply_(dk, rope[0], rope[1], KV, 128, inverse=True)
    if out is not None:
        out[0].copy_(dq); out[1].copy_(dk); out[2].copy_(dv)
        return out
    return dq, dk, dv


_NAMES = ["gemm", "gemm_swiglu", "swiglu_bwd", "rmsnorm_fwd", "rmsnorm_bwd", "rope_table", "rope_apply_", "cast_bf16",
          "embed_add", "AttnPlan", "attn_fwd", "attn_bwd"]


def install():
    """Patch touchnet_b200.ops in THIS process (call inside spawned gloo workers / under monkeypatch)."""
    saved = {n: getattr(ops, n) for n in _NAMES}
    saved["_FUSE_QKV"], saved["_chk"], saved["_FUSE_DSWIGLU"] = ops._FUSE_QKV, ops._chk, ops._FUSE_DSWIGLU
    g = globals()
    for n in _NAMES:
        setattr(ops, n, g[n])
    ops._FUSE_QKV = False                     # the segmented-QKV launches have no stand-in: three plain GEMMs
    ops._FUSE_DSWIGLU = False                 # fused dgrad + SwiGLU backward: the two-op form it is bit-identical to
    ops._chk = lambda *a, **k: None
    return saved


def uninstall(saved):
    for n, f in saved.items():
        setattr(ops, n, f)
