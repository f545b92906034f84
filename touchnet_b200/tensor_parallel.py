"""Tensor parallelism with sequence parallelism for the B200 modules (BASELINE config 5: TP=2 x FSDP2=4).

Same sharding as the reference's plan (ref: touchnet/models/llama/parallelize_llama.py:105-196):

    embed_tokens       rows (vocabulary) split               RowwiseParallel(input Replicate, output Shard(1))
    q/k/v/gate/up      output rows split   -> Shard(0)       ColwiseParallel    (heads H/tp, KV/tp; ffn/tp columns)
    o_proj/down_proj   input columns split -> Shard(1)       RowwiseParallel(output Shard(1))
    every RMSNorm      weight replicated, runs on the rank's T/tp rows      SequenceParallel
    lm_head            rows (vocabulary) split, logits gathered             ColwiseParallel(input Shard(1), output Replicate)

Parameters become DTensors on the tp mesh with exactly those placements, so FSDP2 (applied afterwards over the dp mesh,
ref: touchnet/models/helper_func.py:134-202), DCP checkpoints and the reference's converters see what they expect.  The
reference lets DTensor module hooks insert the collectives around nn.Linear.forward; the B200 decoder block is ONE autograd
node that reads the weights directly, so the collectives are issued explicitly inside it (ops.DecoderLayerFn):

    forward    rmsnorm(T/tp rows) -> all-gather -> QKV GEMM (local heads) -> attention (local heads, whole sequence)
               -> o_proj GEMM (partial sums over tp) -> reduce-scatter -> + residual           (same again for the MLP)
    backward   the mirror image: all-gather where forward reduce-scattered and vice versa; the replicated norm weights'
               gradients (partial sums over sequence shards) are all-reduced.

Between blocks the residual stream lives sequence-sharded: [B, T/tp, d] per rank.  `attention_mask` / `position_ids` /
`input_ids` / `input_features` are the full [B, T] tensors on every tp rank (tp ranks share their batch,
ref: touchnet/utils/distributed.py:116-157 dp coordinates exclude tp).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import ops

BF16 = torch.bfloat16


def is_dtensor(t) -> bool:
    return ops._is_dtensor(t)


def local(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """The rank-local shard of a (possibly) DTensor parameter; gradients flow back with the parameter's placement."""
    if t is not None and is_dtensor(t):
        return t.to_local()
    return t


class TPContext:
    """Collectives of one tp group on [B*rows, C] activations that are sharded on the sequence (rows) dimension."""

    def __init__(self, group: dist.ProcessGroup, B: int):
        self.group = group
        self.size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.B = B

    def ag(self, x: torch.Tensor) -> torch.Tensor:
        """[B*Tl, C] -> [B*T, C]: all-gather along the sequence."""
        xs = x.contiguous()
        out = torch.empty((self.size * xs.shape[0],) + tuple(xs.shape[1:]), dtype=xs.dtype, device=xs.device)
        dist.all_gather_into_tensor(out, xs, group=self.group)
        if self.B == 1:
            return out
        Tl = xs.shape[0] // self.B
        return out.view(self.size, self.B, Tl, -1).permute(1, 0, 2, 3).reshape(self.B * self.size * Tl, -1)

    def rs(self, x: torch.Tensor) -> torch.Tensor:
        """[B*T, C] partial sums -> [B*Tl, C]: reduce-scatter (sum) along the sequence, in the tensor's dtype."""
        rows = x.shape[0]
        if rows % (self.B * self.size) != 0:
            raise ops._lib.TouchNetB200Error(f"sequence of {rows // self.B} rows does not split over tp={self.size}")
        if self.B == 1:
            xs = x.contiguous()
        else:
            Tl = rows // self.B // self.size
            xs = x.view(self.B, self.size, Tl, -1).permute(1, 0, 2, 3).contiguous()
        out = torch.empty((rows // self.size, x.shape[-1]), dtype=x.dtype, device=x.device)
        dist.reduce_scatter_tensor(out, xs.view(rows, -1), op=dist.ReduceOp.SUM, group=self.group)
        return out

    def all_reduce_(self, x: torch.Tensor) -> torch.Tensor:
        dist.all_reduce(x, op=dist.ReduceOp.SUM, group=self.group)
        return x

    def seq_slice(self, x: torch.Tensor) -> torch.Tensor:
        """This rank's rows of a full [B, T, ...] tensor."""
        Tl = x.shape[1] // self.size
        return x[:, self.rank * Tl:(self.rank + 1) * Tl]


class _SeqGather(torch.autograd.Function):
    """all-gather along the sequence; backward reduce-scatters (the consumers' gradients are partial sums)."""

    @staticmethod
    def forward(ctx, x, tp: TPContext):
        ctx.tp = tp
        return tp.ag(x)

    @staticmethod
    def backward(ctx, dy):
        return ctx.tp.rs(dy.contiguous()), None


class _VocabGather(torch.autograd.Function):
    """[rows, V/tp] -> [rows, V] (Shard(-1) -> Replicate).  Every tp rank computes the same loss on the gathered logits,
    so the gradient arrives replicated and the backward is this rank's column slice."""

    @staticmethod
    def forward(ctx, x, tp: TPContext):
        ctx.tp, ctx.vl = tp, x.shape[-1]
        xs = x.contiguous()
        out = torch.empty((tp.size * xs.shape[0], xs.shape[1]), dtype=xs.dtype, device=xs.device)
        dist.all_gather_into_tensor(out, xs, group=tp.group)
        return out.view(tp.size, xs.shape[0], xs.shape[1]).permute(1, 0, 2).reshape(xs.shape[0], tp.size * xs.shape[1])

    @staticmethod
    def backward(ctx, dy):
        tp, vl = ctx.tp, ctx.vl
        return dy[:, tp.rank * vl:(tp.rank + 1) * vl].contiguous(), None


class _ReplicatedParam(torch.autograd.Function):
    """A replicated parameter consumed on sequence shards: its gradient is a partial sum per rank -> all-reduce."""

    @staticmethod
    def forward(ctx, w, group):
        ctx.group = group
        return w.view_as(w)

    @staticmethod
    def backward(ctx, dw):
        dw = dw.contiguous().clone()
        dist.all_reduce(dw, op=dist.ReduceOp.SUM, group=ctx.group)
        return dw, None


def replicated_param(w: torch.Tensor, group) -> torch.Tensor:
    return _ReplicatedParam.apply(local(w), group)


class _VocabParallelEmbed(torch.autograd.Function):
    """embed_tokens with the vocabulary split over tp (ref plan: RowwiseParallel(input Replicate, output Shard(1))):
    local lookup (zeros for ids owned by other ranks) -> reduce-scatter along the sequence."""

    @staticmethod
    def forward(ctx, ids, w_local, tp: TPContext):
        B, T = ids.shape
        vl = w_local.shape[0]
        loc = ids.reshape(-1) - tp.rank * vl
        mine = (loc >= 0) & (loc < vl)
        loc = loc.clamp(0, vl - 1)
        e = w_local.detach()[loc].to(BF16) * mine[:, None].to(BF16)
        ctx.save_for_backward(loc, mine)
        ctx.tp, ctx.w_shape, ctx.w_dtype = tp, w_local.shape, w_local.dtype
        return tp.rs(e).view(B, T // tp.size, -1)

    @staticmethod
    def backward(ctx, de):
        loc, mine = ctx.saved_tensors
        tp = ctx.tp
        de_full = tp.ag(de.reshape(-1, de.shape[-1]).contiguous())
        dw = torch.zeros(ctx.w_shape, dtype=ctx.w_dtype, device=de.device)
        dw.index_add_(0, loc, (de_full * mine[:, None].to(de_full.dtype)).to(ctx.w_dtype))
        return None, dw, None


def embed(ids: torch.Tensor, embed_weight: torch.Tensor, tp: TPContext) -> torch.Tensor:
    """[B, T] ids -> [B, T/tp, d] sequence-sharded embeddings."""
    if ids.shape[1] % tp.size != 0:
        raise ops._lib.TouchNetB200Error(f"T={ids.shape[1]} does not split over tp={tp.size}")
    return _VocabParallelEmbed.apply(ids, local(embed_weight), tp)


def lm_head(h_s: torch.Tensor, weight: torch.Tensor, tp: TPContext) -> torch.Tensor:
    """[B, T/tp, d] -> replicated logits [B, T, V] (ref plan: ColwiseParallel(input Shard(1), output Replicate))."""
    B, Tl, d = h_s.shape
    h = _SeqGather.apply(h_s.reshape(B * Tl, d), tp)
    logits_l = ops.linear(h, local(weight))
    return _VocabGather.apply(logits_l, tp).view(B, Tl * tp.size, -1)


# ---------------------------------------------------------------------------------------------------------------
# sharding the parameters (what `parallelize_module` does in the reference)
# ---------------------------------------------------------------------------------------------------------------
def _distribute(mod: torch.nn.Module, name: str, mesh, placement) -> None:
    from torch.distributed.tensor import distribute_tensor
    p = getattr(mod, name, None)
    if p is None or is_dtensor(p):
        return
    dt = distribute_tensor(p.data, mesh, [placement])
    mod.register_parameter(name, torch.nn.Parameter(dt, requires_grad=p.requires_grad))


def apply_tp(model: torch.nn.Module, tp_mesh) -> torch.nn.Module:
    """Shard a B200LlamaForCausalLM / B200TouchAudioForCausalLM over `tp_mesh` (1-D DeviceMesh) in place.  Call before
    `fully_shard` (ref order: touchnet/models/llama/parallelize_llama.py:43-81).  The audio projector stays replicated
    (it is outside the reference's plan as well) and runs on the rank's sequence shard."""
    from torch.distributed.tensor import Replicate, Shard
    from . import modeling
    lm = model.language_model if hasattr(model, "language_model") else model
    base = lm.model
    tp = tp_mesh.size()
    cfg = base.config
    H = cfg.num_attention_heads
    KV = getattr(cfg, "num_key_value_heads", H) or H
    if H % tp or KV % tp or cfg.intermediate_size % tp or cfg.vocab_size % tp:
        raise ops._lib.TouchNetB200Error(
            f"tp={tp} must divide heads ({H}), kv heads ({KV}), ffn ({cfg.intermediate_size}) and vocab ({cfg.vocab_size})")
    tied = lm.lm_head.weight is base.embed_tokens.weight
    _distribute(base.embed_tokens, "weight", tp_mesh, Shard(0))
    _distribute(base.norm, "weight", tp_mesh, Replicate())
    if tied:
        lm.lm_head.weight = base.embed_tokens.weight
    else:
        _distribute(lm.lm_head, "weight", tp_mesh, Shard(0))
    for layer in base.layers:
        a, m = layer.self_attn, layer.mlp
        for lin in (a.q_proj, a.k_proj, a.v_proj, m.gate_proj, m.up_proj):
            _distribute(lin, "weight", tp_mesh, Shard(0))
            _distribute(lin, "bias", tp_mesh, Shard(0))
        for lin in (a.o_proj, m.down_proj):
            _distribute(lin, "weight", tp_mesh, Shard(1))
        _distribute(layer.input_layernorm, "weight", tp_mesh, Replicate())
        _distribute(layer.post_attention_layernorm, "weight", tp_mesh, Replicate())
    for mod in model.modules():
        if isinstance(mod, (modeling.B200LlamaModel, modeling.B200LlamaForCausalLM, modeling.B200TouchAudioForCausalLM)):
            mod.tp_group = tp_mesh.get_group()
    return model
