"""Context parallelism for the packed attention (BASELINE config 4).

The reference shards every per-token buffer on the sequence dimension over the `cp` mesh
(ref: touchnet/bin/train.py:363-387, touchnet/utils/distributed.py:292-315) and lets torch's experimental
`context_parallel` patch SDPA with an all-gather ("allgather" rotate method, ref: touchnet/bin/__init__.py:307-317) -
which ignores the document mask.  Here the semantics are defined and exact: the block-causal document mask
on the FULL sequence, computed with the sequence sharded contiguously over the cp ranks:

    forward   K/V of all ranks are all-gathered (NCCL over NVLink), each rank runs the packed-attention kernel on its own
              query window against the global K/V (kernels take a query window: Tq rows at block offset q_blk_off);
    backward  each rank produces dQ for its rows and partial dK/dV for ALL rows; a reduce-scatter (sum) returns every
              rank the dK/dV of its own rows.

Everything else on the path (embedding, RMSNorm, GEMMs, RoPE with the sharded position_ids, loss) is token-local and runs
unchanged on the T/cp rows of the rank.  Contiguous (not zig-zag) sharding: with packed documents of bounded length the
work per rank is balanced; a single document spanning the whole sequence would load the last rank most.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import ops


def make_cp_plan(local_doc_ids: torch.Tensor, group: Optional[dist.ProcessGroup], load_balance: bool = False) -> ops.AttnPlan:
    """All-gather the (tiny) per-rank document-id slices into the global [B, T] ids and build the attention plan of this
    rank's query window(s).  local_doc_ids: [B, T/cp] (the `attention_mask` buffer as the reference shards it).

    load_balance=False: rank r holds the contiguous rows [r*T/cp, (r+1)*T/cp).
    load_balance=True : torch's head-tail layout (the default of `context_parallel`, _HeadTailLoadBalancer): the sequence
                        is cut into 2*cp chunks and rank r holds chunk r followed by chunk 2*cp-1-r, so that every rank
                        gets the same causal work; the attention kernels then run once per chunk (two query windows)."""
    cp = dist.get_world_size(group)
    rank = dist.get_rank(group)
    B, Tl = local_doc_ids.shape
    unit = 256 if load_balance else 128
    if Tl % unit != 0:
        raise ops._lib.TouchNetB200Error(f"context parallelism needs T/cp to be a multiple of {unit}, got {Tl}")
    ids = local_doc_ids.to(torch.int32).contiguous()
    parts = [torch.empty_like(ids) for _ in range(cp)]
    dist.all_gather(parts, ids, group=group)
    if not load_balance:
        full = torch.cat(parts, dim=1)
        plan = ops.AttnPlan(full, Tq=Tl, q_blk_off=rank * Tl // 128)
        plan.cp_group = group
        return plan
    Tw = Tl // 2
    full = _head_tail_to_global(torch.stack(parts).unsqueeze(-1), B).view(B, cp * Tl)
    plan = ops.AttnPlan(full, Tq=Tw, q_blk_off=rank * Tw // 128)
    plan.cp_group = group
    plan.cp_windows = [rank * Tw // 128, (2 * cp - 1 - rank) * Tw // 128]      # first block of each local chunk
    return plan


def _head_tail_to_global(g: torch.Tensor, B: int) -> torch.Tensor:
    """[cp, B, Tl, C] per-rank head-tail shards -> [B*T, C] in global sequence order."""
    cp, _, Tl, C = g.shape
    Tw = Tl // 2
    out = torch.empty((B, 2 * cp, Tw, C), dtype=g.dtype, device=g.device)
    gv = g.view(cp, B, 2, Tw, C)
    for r in range(cp):
        out[:, r] = gv[r, :, 0]
        out[:, 2 * cp - 1 - r] = gv[r, :, 1]
    return out.view(B * 2 * cp * Tw, C)


def _global_to_head_tail(x: torch.Tensor, B: int, cp: int) -> torch.Tensor:
    """[B*T, C] global order -> [cp, B, Tl, C]: slot r = what rank r holds (chunk r, then chunk 2*cp-1-r)."""
    C = x.shape[-1]
    Tw = x.shape[0] // B // (2 * cp)
    xv = x.view(B, 2 * cp, Tw, C)
    out = torch.empty((cp, B, 2, Tw, C), dtype=x.dtype, device=x.device)
    for r in range(cp):
        out[r, :, 0] = xv[:, r]
        out[r, :, 1] = xv[:, 2 * cp - 1 - r]
    return out.view(cp, B, 2 * Tw, C)


def _gather_seq(x: torch.Tensor, B: int, group, head_tail: bool = False) -> torch.Tensor:
    """[B*Tl, C] local rows -> [B*T, C] global rows (sequence-contiguous per batch row)."""
    cp = dist.get_world_size(group)
    Tl = x.shape[0] // B
    xs = x.contiguous()
    out = torch.empty((cp * xs.shape[0],) + tuple(xs.shape[1:]), dtype=xs.dtype, device=xs.device)
    dist.all_gather_into_tensor(out, xs, group=group)
    if head_tail:
        return _head_tail_to_global(out.view(cp, B, Tl, -1), B)
    return out.view(cp, B, Tl, -1).permute(1, 0, 2, 3).reshape(B * cp * Tl, -1)


def _reduce_scatter_seq(x: torch.Tensor, B: int, group, head_tail: bool = False) -> torch.Tensor:
    """[B*T, C] partial sums over global rows -> [B*Tl, C] summed rows of this rank (fp32 accumulation on the wire)."""
    cp = dist.get_world_size(group)
    T = x.shape[0] // B
    Tl = T // cp
    if head_tail:
        xs = _global_to_head_tail(x, B, cp).float().view(cp * B * Tl, -1)
    else:
        xs = x.view(B, cp, Tl, -1).permute(1, 0, 2, 3).contiguous().float().view(cp * B * Tl, -1)
    out = torch.empty((B * Tl, xs.shape[-1]), dtype=torch.float32, device=x.device)
    dist.reduce_scatter_tensor(out, xs, op=dist.ReduceOp.SUM, group=group)
    return out.view(B * Tl, -1).to(x.dtype)


# ---------------------------------------------------------------------------------------------------------------
# halo exchange (default; TN_CP_HALO=0 = whole-shard all-gather): move only the K/V rows a rank's queries can reach
# ---------------------------------------------------------------------------------------------------------------
# With packed documents a query never looks past the start of its own document, so rank r needs K/V only from the first
# block of the earliest document that reaches into its window (the per-block `kv_lo` the attention kernels already use)
# up to its own rows - typically a fraction of ONE neighbour's shard instead of every other rank's whole shard (SURVEY
# 8(e): "halo exchange - an optimisation the reference does not do").  Same for dK/dV on the way back.  The exchanged
# rows are whole 128-row blocks, so every block the kernels load is fully initialised.
def _halo_enabled() -> bool:
    """Default ON (validated on 2 x B200: same logits / gradients as the all-gather form, tools/check_cp.py, and faster);
    TN_CP_HALO=0 selects the reference-style whole-shard K/V all-gather."""
    import os
    return os.environ.get("TN_CP_HALO", "1") != "0"


def halo_first_blocks(plan: ops.AttnPlan) -> list:
    """First K/V block each cp rank needs (<= its own first block), from the block metadata every rank holds.
    One small device->host read per step, cached on the plan."""
    cached = getattr(plan, "_halo_lo", None)
    if cached is not None:
        return cached
    B, nblk = plan.B, (plan.T + 127) // 128
    nq = plan.Tq // 128
    cp = nblk // nq
    m = plan.meta[: B * nblk * 4].view(B, nblk, 4)
    kv_lo, kv_end = m[..., 0], m[..., 1]
    lo = torch.where(kv_end > kv_lo, kv_lo, torch.full_like(kv_lo, nblk))      # padding-only blocks load nothing
    lo = lo.view(B, cp, nq).amin(dim=(0, 2))
    own = torch.arange(cp, device=lo.device, dtype=lo.dtype) * nq
    plan._halo_lo = torch.minimum(lo, own).tolist()
    return plan._halo_lo


def _halo_pairs(plan: ops.AttnPlan):
    """[(owner s, consumer r, first row, end row)]: rows of rank s that rank r > s needs (global row indices)."""
    lo = halo_first_blocks(plan)
    Tl = plan.Tq
    out = []
    for r in range(len(lo)):
        for s in range(r):
            a, b = max(lo[r] * 128, s * Tl), (s + 1) * Tl
            if a < b:
                out.append((s, r, a, b))
    return out


def _halo_gather(x: torch.Tensor, plan: ops.AttnPlan) -> torch.Tensor:
    """[B*Tl, C] local rows -> [B*T, C] with this rank's rows and the halo rows it needs filled in (the rest of the
    buffer is never read by the kernels)."""
    group = plan.cp_group
    me = dist.get_rank(group)
    B, T, Tl = plan.B, plan.T, plan.Tq
    C = x.shape[1]
    full = torch.empty((B, T, C), dtype=x.dtype, device=x.device)
    xl = x.view(B, Tl, C)
    full[:, me * Tl:(me + 1) * Tl] = xl
    reqs, recvs = [], []
    for s, r, a, b in _halo_pairs(plan):
        if me == s:
            reqs.append(dist.P2POp(dist.isend, xl[:, a - s * Tl:b - s * Tl].contiguous(), dist.get_global_rank(group, r), group))
        elif me == r:
            buf = torch.empty((B, b - a, C), dtype=x.dtype, device=x.device)
            recvs.append((buf, a, b))
            reqs.append(dist.P2POp(dist.irecv, buf, dist.get_global_rank(group, s), group))
    if reqs:
        for w in dist.batch_isend_irecv(reqs):
            w.wait()
    for buf, a, b in recvs:
        full[:, a:b] = buf
    return full.view(B * T, C)


def _halo_reduce(d_full: torch.Tensor, plan: ops.AttnPlan) -> torch.Tensor:
    """[B*T, C] partial dK|dV (zero outside the rows this rank's queries reach) -> [B*Tl, C] summed rows of this rank:
    halo rows go back to their owners (fp32 accumulation at the owner)."""
    group = plan.cp_group
    me = dist.get_rank(group)
    B, T, Tl = plan.B, plan.T, plan.Tq
    C = d_full.shape[1]
    df = d_full.view(B, T, C)
    reqs, recvs = [], []
    for s, r, a, b in _halo_pairs(plan):
        if me == r:
            reqs.append(dist.P2POp(dist.isend, df[:, a:b].contiguous(), dist.get_global_rank(group, s), group))
        elif me == s:
            buf = torch.empty((B, b - a, C), dtype=d_full.dtype, device=d_full.device)
            recvs.append((buf, a - s * Tl, b - s * Tl))
            reqs.append(dist.P2POp(dist.irecv, buf, dist.get_global_rank(group, r), group))
    if reqs:
        for w in dist.batch_isend_irecv(reqs):
            w.wait()
    own = df[:, me * Tl:(me + 1) * Tl]
    if not recvs:
        return own.reshape(B * Tl, C).contiguous()
    acc = own.float()
    for buf, a, b in recvs:
        acc[:, a:b] += buf.float()
    return acc.to(d_full.dtype).view(B * Tl, C)


def _window_plan(plan: ops.AttnPlan, q_blk_off: int) -> ops.AttnPlan:
    import copy
    pw = copy.copy(plan)                  # shares the document ids / block metadata, differs in the query window
    pw.q_blk_off = q_blk_off
    return pw


def _window_rows(x: torch.Tensor, B: int, w: int) -> torch.Tensor:
    """Rows of local chunk w (0 = head, 1 = tail) of a [B*2*Tw, C] tensor as a dense [B*Tw, C] tensor."""
    C = x.shape[-1]
    Tw = x.shape[0] // B // 2
    xw = x.view(B, 2, Tw, C)[:, w]
    return xw.reshape(B * Tw, C) if B == 1 else xw.contiguous().view(B * Tw, C)


def _cp_attn_fwd_head_tail(q, k, v, plan, H, KV, scale):
    B = plan.B
    kf = _gather_seq(k, B, plan.cp_group, head_tail=True)
    vf = _gather_seq(v, B, plan.cp_group, head_tail=True)
    Tw = plan.Tq
    o = torch.empty((B, 2, Tw, q.shape[1]), dtype=q.dtype, device=q.device)
    lses = []
    for w, off in enumerate(plan.cp_windows):
        o_w, lse_w = ops.attn_fwd(_window_rows(q, B, w), kf, vf, _window_plan(plan, off), H, KV, scale)
        o[:, w] = o_w.view(B, Tw, -1)
        lses.append(lse_w)
    return o.view(B * 2 * Tw, -1), torch.cat(lses, dim=-1), kf, vf


def _cp_attn_bwd_head_tail(q, kf, vf, o, do, lse, plan, H, KV, scale):
    B, Tw = plan.B, plan.Tq
    dq = torch.empty((B, 2, Tw, q.shape[1]), dtype=q.dtype, device=q.device)
    dk_full = dv_full = None
    for w, off in enumerate(plan.cp_windows):
        lse_w = lse[..., w * Tw:(w + 1) * Tw].contiguous()
        dq_w, dk_w, dv_w = ops.attn_bwd(_window_rows(q, B, w), kf, vf, _window_rows(o, B, w), _window_rows(do, B, w),
                                        lse_w, _window_plan(plan, off), H, KV, scale)
        dq[:, w] = dq_w.view(B, Tw, -1)
        dk_full = dk_w.float() if dk_full is None else dk_full + dk_w.float()
        dv_full = dv_w.float() if dv_full is None else dv_full + dv_w.float()
    dk = _reduce_scatter_seq(dk_full, B, plan.cp_group, head_tail=True)
    dv = _reduce_scatter_seq(dv_full, B, plan.cp_group, head_tail=True)
    return dq.view(B * 2 * Tw, -1), dk.to(q.dtype), dv.to(q.dtype)


def cp_attn_fwd(q, k, v, plan: ops.AttnPlan, H: int, KV: int, scale: float):
    """q [B*Tl, H*128] (already rotated), k/v [B*Tl, KV*128] local -> (o local, lse local, k_full, v_full)."""
    if getattr(plan, "cp_windows", None) is not None:
        return _cp_attn_fwd_head_tail(q, k, v, plan, H, KV, scale)
    if _halo_enabled():
        kv = _halo_gather(torch.cat([k, v], dim=1), plan)          # one exchange for both; views keep the row stride
        c = k.shape[1]
        kf, vf = kv[:, :c], kv[:, c:]
    else:
        kf = _gather_seq(k, plan.B, plan.cp_group)
        vf = _gather_seq(v, plan.B, plan.cp_group)
    o, lse = ops.attn_fwd(q, kf, vf, plan, H, KV, scale)
    return o, lse, kf, vf


def cp_attn_bwd(q, kf, vf, o, do, lse, plan: ops.AttnPlan, H: int, KV: int, scale: float):
    if getattr(plan, "cp_windows", None) is not None:
        return _cp_attn_bwd_head_tail(q, kf, vf, o, do, lse, plan, H, KV, scale)
    dq, dk_full, dv_full = ops.attn_bwd(q, kf, vf, o, do, lse, plan, H, KV, scale)
    if _halo_enabled():
        c = dk_full.shape[1]
        d = _halo_reduce(torch.cat([dk_full, dv_full], dim=1), plan)
        return dq, d[:, :c].contiguous(), d[:, c:].contiguous()
    dk = _reduce_scatter_seq(dk_full, plan.B, plan.cp_group)
    dv = _reduce_scatter_seq(dv_full, plan.B, plan.cp_group)
    return dq, dk, dv


def enable_context_parallel(model: torch.nn.Module, group: Optional[dist.ProcessGroup], load_balance: bool = False) -> None:
    """Mark a B200LlamaForCausalLM / B200TouchAudioForCausalLM as running on sequence shards of the given cp group
    (pass None to switch it off).  The caller feeds every per-token buffer already sharded on dim 1, as the reference's
    `create_context_parallel_ctx` does (ref: touchnet/bin/train.py:363-387): contiguous shards by default,
    load_balance=True for torch's head-tail layout (what `context_parallel` produces unless its load balancing is
    switched off)."""
    from . import modeling
    for m in model.modules():
        if isinstance(m, modeling.B200LlamaModel):
            m.cp_group = group
            m.cp_load_balance = bool(load_balance)
