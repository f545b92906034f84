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


def make_cp_plan(local_doc_ids: torch.Tensor, group: Optional[dist.ProcessGroup]) -> ops.AttnPlan:
    """All-gather the (tiny) per-rank document-id slices into the global [B, T] ids and build the attention plan of this
    rank's query window.  local_doc_ids: [B, T/cp] (the `attention_mask` buffer as the reference shards it)."""
    cp = dist.get_world_size(group)
    rank = dist.get_rank(group)
    B, Tl = local_doc_ids.shape
    if Tl % 128 != 0:
        raise ops._lib.TouchNetB200Error(f"context parallelism needs T/cp to be a multiple of 128, got {Tl}")
    ids = local_doc_ids.to(torch.int32).contiguous()
    parts = [torch.empty_like(ids) for _ in range(cp)]
    dist.all_gather(parts, ids, group=group)
    full = torch.cat(parts, dim=1)
    plan = ops.AttnPlan(full, Tq=Tl, q_blk_off=rank * Tl // 128)
    plan.cp_group = group
    return plan


def _gather_seq(x: torch.Tensor, B: int, group) -> torch.Tensor:
    """[B*Tl, C] local rows -> [B*T, C] global rows (sequence-contiguous per batch row)."""
    cp = dist.get_world_size(group)
    Tl = x.shape[0] // B
    xs = x.contiguous()
    out = torch.empty((cp * xs.shape[0],) + tuple(xs.shape[1:]), dtype=xs.dtype, device=xs.device)
    dist.all_gather_into_tensor(out, xs, group=group)
    return out.view(cp, B, Tl, -1).permute(1, 0, 2, 3).reshape(B * cp * Tl, -1)


def _reduce_scatter_seq(x: torch.Tensor, B: int, group) -> torch.Tensor:
    """[B*T, C] partial sums over global rows -> [B*Tl, C] summed rows of this rank (fp32 accumulation on the wire)."""
    cp = dist.get_world_size(group)
    T = x.shape[0] // B
    Tl = T // cp
    xs = x.view(B, cp, Tl, -1).permute(1, 0, 2, 3).contiguous().float().view(cp * B * Tl, -1)
    out = torch.empty((B * Tl, xs.shape[-1]), dtype=torch.float32, device=x.device)
    dist.reduce_scatter_tensor(out, xs, op=dist.ReduceOp.SUM, group=group)
    return out.view(B * Tl, -1).to(x.dtype)


def cp_attn_fwd(q, k, v, plan: ops.AttnPlan, H: int, KV: int, scale: float):
    """q [B*Tl, H*128] (already rotated), k/v [B*Tl, KV*128] local -> (o local, lse local, k_full, v_full)."""
    kf = _gather_seq(k, plan.B, plan.cp_group)
    vf = _gather_seq(v, plan.B, plan.cp_group)
    o, lse = ops.attn_fwd(q, kf, vf, plan, H, KV, scale)
    return o, lse, kf, vf


def cp_attn_bwd(q, kf, vf, o, do, lse, plan: ops.AttnPlan, H: int, KV: int, scale: float):
    dq, dk_full, dv_full = ops.attn_bwd(q, kf, vf, o, do, lse, plan, H, KV, scale)
    dk = _reduce_scatter_seq(dk_full, plan.B, plan.cp_group)
    dv = _reduce_scatter_seq(dv_full, plan.B, plan.cp_group)
    return dq, dk, dv


def enable_context_parallel(model: torch.nn.Module, group: Optional[dist.ProcessGroup]) -> None:
    """Mark a B200LlamaForCausalLM / B200TouchAudioForCausalLM as running on sequence shards of the given cp group
    (pass None to switch it off).  The caller feeds every per-token buffer already sharded on dim 1, as the reference's
    `create_context_parallel_ctx` does (ref: touchnet/bin/train.py:363-387)."""
    from . import modeling
    for m in model.modules():
        if isinstance(m, modeling.B200LlamaModel):
            m.cp_group = group
