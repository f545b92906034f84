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

    # -- what the decoder block calls (ops.DecoderLayerFn); PeerTPContext overrides these three --------------------
    def begin_step(self) -> None:
        """Called once per model forward, before the first block."""

    def gather_rows(self, x: torch.Tensor, keep: bool = False) -> torch.Tensor:
        """[B*Tl, C] -> [B*T, C].  keep=True: the result is saved for backward by the caller."""
        return self.ag(x)

    def reduce_scatter(self, partial: torch.Tensor) -> torch.Tensor:
        """[B*T, C] partial sums -> [B*Tl, C] summed rows of this rank."""
        return self.rs(partial)

    def gemm_reduce_scatter(self, a: torch.Tensor, w: torch.Tensor, residual_s: torch.Tensor) -> torch.Tensor:
        """reduce_scatter(a · wᵀ) + residual_s: the row-parallel projections (o_proj, down_proj) of the forward."""
        return self.rs(ops.gemm(a, w)).add_(residual_s)

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
    def forward(ctx, ids, w_local, tp: TPContext, padding_idx=None):
        B, T = ids.shape
        vl = w_local.shape[0]
        loc = ids.reshape(-1) - tp.rank * vl
        mine = (loc >= 0) & (loc < vl)
        loc = loc.clamp(0, vl - 1)
        e = w_local.detach()[loc].to(BF16) * mine[:, None].to(BF16)
        if padding_idx is not None:     # nn.Embedding(padding_idx=...): the pad row receives no gradient
            mine = mine & (ids.reshape(-1) != padding_idx)
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
        return None, dw, None, None


def embed(ids: torch.Tensor, embed_weight: torch.Tensor, tp: TPContext, padding_idx=None) -> torch.Tensor:
    """[B, T] ids -> [B, T/tp, d] sequence-sharded embeddings."""
    if ids.shape[1] % tp.size != 0:
        raise ops._lib.TouchNetB200Error(f"T={ids.shape[1]} does not split over tp={tp.size}")
    return _VocabParallelEmbed.apply(ids, local(embed_weight), tp, padding_idx)


def lm_head(h_s: torch.Tensor, weight: torch.Tensor, tp: TPContext) -> torch.Tensor:
    """[B, T/tp, d] -> replicated logits [B, T, V] (ref plan: ColwiseParallel(input Shard(1), output Replicate))."""
    B, Tl, d = h_s.shape
    h = _SeqGather.apply(h_s.reshape(B * Tl, d), tp)
    logits_l = ops.linear(h, local(weight))
    return _VocabGather.apply(logits_l, tp).view(B, Tl * tp.size, -1)


def lm_head_loss_parallel(h_s: torch.Tensor, weight: torch.Tensor, tp: TPContext):
    """Loss-parallel form (ref: touchnet/models/llama/parallelize_llama.py:127-131 `lm_head` output Shard(-1) +
    touchnet/utils/distributed.py:322-323 `loss_parallel()`): the logits stay sharded on the vocabulary - this rank's
    [B, T, V/tp] columns - inside a handle that loss.cross_entropy_loss / loss.accuracy consume; only per-row statistics
    (logsumexp pieces, label logit, best value) cross the tp group instead of the [B*T, V] tensor."""
    from . import loss as _loss
    B, Tl, d = h_s.shape
    h = _SeqGather.apply(h_s.reshape(B * Tl, d), tp)
    logits_l = ops.linear(h, local(weight))
    vl = logits_l.shape[-1]
    return _loss.VocabParallelLogits(logits_l.view(B, Tl * tp.size, vl), tp.group, tp.rank * vl, vl * tp.size)


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


# ---------------------------------------------------------------------------------------------------------------
# NCCL-free variant: kernels store straight into the peers' memory (TN_TP_PEER=1; what bench.py --tp runs)
# ---------------------------------------------------------------------------------------------------------------
class SymmPeerMemory:
    """Symmetric buffers of one tp group through torch.distributed._symmetric_memory (CUDA IPC / NVLink P2P):
    `alloc` returns, for every rank of the group, a tensor aliasing THAT rank's buffer, so a kernel whose output pointer
    is `views[peer]` writes over NVLink into the peer's HBM."""

    def __init__(self, group: dist.ProcessGroup, device: torch.device):
        import torch.distributed._symmetric_memory as symm_mem
        self._symm = symm_mem
        self.group, self.device = group, device
        self.size = dist.get_world_size(group)
        self._handles = []

    def alloc(self, shape, dtype) -> list:
        t = self._symm.empty(*shape, dtype=dtype, device=self.device)
        h = self._symm.rendezvous(t, self.group)
        self._handles.append((t, h))
        return [h.get_buffer(r, tuple(shape), dtype) for r in range(self.size)]

    def barrier(self, channel: int = 0) -> None:
        """Device-side barrier on the current stream: stores issued before it by any rank are visible to all after it.
        Collectives that may be in flight on different streams at the same time (FSDP2's all-gather and reduce-scatter
        streams) must use different `channel`s of the signal pad."""
        self._handles[0][1].barrier(channel=channel)


class PeerTPContext(TPContext):
    """The block's tp collectives without NCCL (SURVEY 8(e) B200 note: overlap / fusion over peer memory):

      gemm_reduce_scatter   the row-parallel GEMM is launched once per destination rank on that rank's rows of A; its
                            epilogue (TMA store) writes the partial tile directly into the destination's receive slot
                            over NVLink while later tiles are still being multiplied - the transfer IS the epilogue.
                            After one barrier every rank adds its tp receive slots (+ the residual, fused into the local
                            launch).
      gather_rows           each rank stores its rows into every rank's full buffer, one barrier.
      reduce_scatter        (backward) peers' rows of the partial are copied into their receive slots, barrier, add.

    Receive slots alternate between two buffers so that one barrier per collective suffices (a slot is rewritten only
    after a later barrier, which its reader reaches after it has consumed the slot).  Buffers returned with keep=True are
    per call site and live until the same call site of the next step (they are saved for backward), so activation
    checkpointing is not supported with this context.

    STATUS (round 2): validated on 2 x B200 (tools/check_tp.py: same logits / gradients as the NCCL form) and measured
    (T=16384, 2 GPUs: 40.4 k -> 41.1 k tokens/s; with loss parallel 43.4 k); the wiring is also tested on CPU against the
    unsharded model with shared-memory files standing in for symmetric memory (tests/test_parallel_gloo.py).  Opt-in in the
    library (TN_TP_PEER=1) because buffers kept for backward are per call site (no activation checkpointing)."""

    def __init__(self, group: dist.ProcessGroup, B: int, mem):
        super().__init__(group, B)
        self.mem = mem
        self._kept: dict = {}       # (call index, shape, dtype) -> views of the per-call-site gather buffers
        self._pool: dict = {}       # (tag, shape, dtype) -> [views_parity0, views_parity1]
        self._parity: dict = {}
        self._calls = 0

    def begin_step(self) -> None:
        self._calls = 0

    def _transient(self, tag, shape, dtype):
        key = (tag, tuple(shape), dtype)
        if key not in self._pool:
            self._pool[key] = [self.mem.alloc(shape, dtype), self.mem.alloc(shape, dtype)]
            self._parity[key] = 0
        par = self._parity[key]
        self._parity[key] = par ^ 1
        return self._pool[key][par]

    def _rows(self, full: torch.Tensor, r: int) -> torch.Tensor:
        """Rank r's rows of a full [B*T, C] tensor as a [B, Tl, C] view."""
        C = full.shape[-1]
        return full.view(self.B, self.size, -1, C)[:, r]

    def gather_rows(self, x: torch.Tensor, keep: bool = False) -> torch.Tensor:
        rows, C = x.shape
        shape = (rows * self.size, C)
        if keep:
            key = (self._calls, shape, x.dtype)
            self._calls += 1
            if key not in self._kept:
                self._kept[key] = self.mem.alloc(shape, x.dtype)
            views = self._kept[key]
        else:
            views = self._transient("gather", shape, x.dtype)
        xs = x.view(self.B, rows // self.B, C)
        for r in range(self.size):                      # own copy included: every rank ends up with all rows
            self._rows(views[r], self.rank).copy_(xs)
        self.mem.barrier()
        return views[self.rank]

    def _sum_slots(self, slots: torch.Tensor) -> torch.Tensor:
        out = slots[0] + slots[1]
        for r in range(2, self.size):
            out += slots[r]
        return out

    def reduce_scatter(self, partial: torch.Tensor) -> torch.Tensor:
        rows, C = partial.shape
        Tl = rows // self.B // self.size
        views = self._transient("rs", (self.size, self.B * Tl, C), partial.dtype)      # [source rank, local rows, C]
        for r in range(self.size):
            views[r][self.rank].view(self.B, Tl, C).copy_(self._rows(partial, r))
        self.mem.barrier()
        return self._sum_slots(views[self.rank])

    def gemm_reduce_scatter(self, a: torch.Tensor, w: torch.Tensor, residual_s: torch.Tensor) -> torch.Tensor:
        rows, K = a.shape
        N = w.shape[0]
        Tl = rows // self.B // self.size
        views = self._transient("gemm_rs", (self.size, self.B * Tl, N), BF16)
        res = residual_s.view(self.B, Tl, N)
        # peers first: their tiles travel over NVLink while the local launch (which also adds the residual) computes
        order = [r for r in range(self.size) if r != self.rank] + [self.rank]
        for r in order:
            a_r = self._rows(a, r)
            slot = views[r][self.rank].view(self.B, Tl, N)
            for b in range(self.B):
                ops.gemm(a_r[b], w, out=slot[b], residual=res[b] if r == self.rank else None)
        self.mem.barrier()
        return self._sum_slots(views[self.rank])


def make_context(group: dist.ProcessGroup, B: int, device: torch.device, cache_on=None) -> TPContext:
    """TPContext of a model forward.  TN_TP_PEER=1 selects the NCCL-free PeerTPContext (cached on `cache_on`, it owns
    symmetric buffers)."""
    import os
    if os.environ.get("TN_TP_PEER", "0") == "0":
        return TPContext(group, B)
    ctx = getattr(cache_on, "_tn_tp_peer_ctx", None) if cache_on is not None else None
    if ctx is None or ctx.B != B or ctx.group is not group:
        ctx = PeerTPContext(group, B, SymmPeerMemory(group, device))
        if cache_on is not None:
            cache_on._tn_tp_peer_ctx = ctx
    return ctx
