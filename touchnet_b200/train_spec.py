"""Plugin registration: how the B200 modules enter the reference's train loop with no edit to touchnet/bin/train.py.

Two seams:

1. Outer plugin API - `TrainSpec` (ref: touchnet/utils/train_spec.py:25-62).  `register()` clones the reference's own
   "llama" / "touch_audio" specs (ref: touchnet/__init__.py:35-117) and swaps `model_cls` for the B200 modules under the
   names "llama_b200" / "touch_audio_b200"; `parallelize_fn` is wrapped (parallelize.py: tensor / context parallelism
   are applied to the fused block, FSDP2 / AC stay the reference's); `loss_fn` / `acc_fn` become the CUDA pack-loss
   cross-entropy and its argmax accuracy (loss.py; same signatures and return values as
   ref: touchnet/loss/cross_entropy.py:12-50 and touchnet/utils/metrics.py:26-50, consumed at train.py:447-450); every
   other callable (dataloader, tokenizer, optimizer, lr schedule, flops ...) stays the reference's.  `--training_model_name touch_audio_b200` then selects it (ref: touchnet/bin/train.py:119).

2. Inner operator API - HF's attention-interface registry
   (`ALL_ATTENTION_FUNCTIONS[config._attn_implementation]`, hf: models/llama/modeling_llama.py:272-286;
   FlexAttention entry hf: integrations/flex_attention.py:262-364).  `register_hf_attention()` installs
   "touchnet_b200": same signature and return convention as `flex_attention_forward`, document ids supplied either as
   the `[B,T]` integer `attention_mask` itself or through `packed_document_ids(...)`.
"""
from __future__ import annotations

import contextlib
import dataclasses
import threading
from dataclasses import dataclass
from typing import Any, Callable, Optional

import torch

from . import loss as _loss
from . import modeling, ops, parallelize

_tls = threading.local()


@dataclass
class TrainSpec:
    """Field-for-field mirror of ref: touchnet/utils/train_spec.py:25-44 (used when the reference is not importable)."""
    name: str
    model_cls: Any
    config_cls: Any
    parallelize_fn: Optional[Callable] = None
    pipelining_fn: Optional[Callable] = None
    build_optimizers_fn: Optional[Callable] = None
    build_lr_schedulers_fn: Optional[Callable] = None
    build_dataloader_fn: Optional[Callable] = None
    build_tokenizer_fn: Optional[Callable] = None
    loss_fn: Optional[Callable] = None
    acc_fn: Optional[Callable] = None
    additional_pre_init_fn: Optional[Callable] = None
    additional_post_init_fn: Optional[Callable] = None
    get_num_flop_per_token_fn: Optional[Callable] = None
    get_num_params_fn: Optional[Callable] = None
    build_metrics_processor_fn: Optional[Callable] = None


_local_specs: dict[str, TrainSpec] = {}


class B200LlamaForCausalLMFused(modeling.B200LlamaForCausalLM):
    """model_cls of the "llama_b200" spec: the fused lm_head + loss is on (the spec's loss_fn / acc_fn consume the
    LazyLogits handle; ref contract: touchnet/bin/train.py:439-450)."""
    fused_linear_ce = True


class B200TouchAudioForCausalLMFused(modeling.B200TouchAudioForCausalLM):
    """model_cls of the "touch_audio_b200" spec (see B200LlamaForCausalLMFused)."""

    def __init__(self, config):
        super().__init__(config)
        self.language_model.fused_linear_ce = True


def get_num_flop_per_token(num_params: int, model_config, seq_len: int) -> int:
    """ref: touchnet/models/llama/__init__.py:39-54 (dense-attention convention, no recompute credit)."""
    l, h = model_config.num_hidden_layers, model_config.num_attention_heads
    q = model_config.hidden_size // model_config.num_attention_heads
    return 6 * num_params + 12 * l * h * q * seq_len


def get_num_params(model: torch.nn.Module, exclude_embedding: bool = False) -> int:
    """ref: touchnet/models/llama/__init__.py:57-67."""
    n = sum(p.numel() for p in model.parameters())
    if exclude_embedding:
        sub = getattr(model, getattr(model, "base_model_prefix", "model"))
        sub = getattr(sub, "model", sub) if not any(isinstance(m, torch.nn.Embedding) for m in sub.children()) else sub
        n -= sum(sum(p.numel() for p in m.parameters()) for m in sub.children() if isinstance(m, torch.nn.Embedding))
    return n


def register(touchnet_pkg=None) -> list[str]:
    """Register "llama_b200" and "touch_audio_b200".  With the reference importable its registry is used (and its own
    specs cloned); otherwise the specs land in this module's registry with the reference-independent callables."""
    names = []
    pairs = (("llama", B200LlamaForCausalLMFused), ("touch_audio", B200TouchAudioForCausalLMFused))
    try:
        if touchnet_pkg is None:
            import touchnet as touchnet_pkg  # type: ignore  # noqa: F401
        from touchnet.utils.train_spec import get_train_spec, register_train_spec  # type: ignore
    except ImportError:
        # the reference package is not importable here: keep the specs in this module's registry.
        # Only ImportError is tolerated - any other failure of the reference branch is a real integration error.
        for base, cls in pairs:
            spec = TrainSpec(name=base + "_b200", model_cls=cls, config_cls=None,
                             parallelize_fn=parallelize.make_parallelize_fn(None),
                             loss_fn=_loss.cross_entropy_loss, acc_fn=_loss.accuracy,
                             get_num_flop_per_token_fn=get_num_flop_per_token, get_num_params_fn=get_num_params)
            _local_specs[spec.name] = spec
            names.append(spec.name)
        return names
    for base, cls in pairs:
        ref_spec = get_train_spec(base)
        spec = dataclasses.replace(ref_spec, name=base + "_b200", model_cls=cls,
                                   parallelize_fn=parallelize.make_parallelize_fn(ref_spec.parallelize_fn),
                                   loss_fn=_loss.cross_entropy_loss, acc_fn=_loss.accuracy)
        try:
            register_train_spec(spec)
        except ValueError:
            pass  # already registered
        names.append(spec.name)
    return names


def get_train_spec(name: str) -> TrainSpec:
    if name not in _local_specs:
        raise ValueError(f"Model {name} is not registered.")
    return _local_specs[name]


# ---------------------------------------------------------------------------------------------------------------
# HF attention-interface seam
# ---------------------------------------------------------------------------------------------------------------
@contextlib.contextmanager
def packed_document_ids(doc_ids: torch.Tensor):
    """Make the `[B,T]` document-id tensor of the current packed batch visible to `hf_attention_forward` when the
    caller's mask plumbing has already turned `attention_mask` into something else (BlockMask / 4-D mask)."""
    prev = getattr(_tls, "plan", None)
    _tls.plan = ops.AttnPlan(doc_ids)
    try:
        yield
    finally:
        _tls.plan = prev


def hf_attention_forward(module, query, key, value, attention_mask=None, dropout: float = 0.0, scaling=None,
                         **kwargs):
    """Drop-in for `flex_attention_forward(module, query [B,H,T,hd], key [B,KV,T,hd], value, attention_mask, ...)`
    (hf: integrations/flex_attention.py:262-364): returns (attn_output [B,T,H,hd], None)."""
    B, H, T, hd = query.shape
    KV = key.shape[1]
    if hd != 128:
        raise ops._lib.TouchNetB200Error(f"touchnet_b200 attention needs head_dim 128, got {hd}")
    if dropout:
        raise ops._lib.TouchNetB200Error("attention dropout is not supported (0.0 in every reference config)")
    if isinstance(attention_mask, torch.Tensor) and attention_mask.dim() == 2 and not attention_mask.is_floating_point():
        plan = ops.AttnPlan(attention_mask)
    else:
        plan = getattr(_tls, "plan", None)
        if plan is None:
            plan = ops.AttnPlan(torch.ones((B, T), dtype=torch.int32, device=query.device))   # plain causal
    q2 = query.transpose(1, 2).reshape(B * T, H * hd)      # a view when q is the usual transposed projection output
    k2 = key.transpose(1, 2).reshape(B * T, KV * hd)
    v2 = value.transpose(1, 2).reshape(B * T, KV * hd)
    dt = query.dtype
    if dt != torch.bfloat16:
        q2, k2, v2 = q2.bfloat16(), k2.bfloat16(), v2.bfloat16()
    scale = float(scaling) if scaling is not None else hd ** -0.5
    o = ops.AttentionFn.apply(q2, k2, v2, plan, H, KV, scale)
    return o.view(B, T, H, hd).to(dt), None


def register_hf_attention(name: str = "touchnet_b200") -> bool:
    """`"attn_implementation": "touchnet_b200"` in the model JSON then selects the kernel (cf. "flex_attention" at
    ref: examples/text/pretrain/fineweb-edu/config/Llama-3_2-1B.json:7)."""
    try:
        from transformers.modeling_utils import ALL_ATTENTION_FUNCTIONS
    except Exception:
        return False
    try:
        ALL_ATTENTION_FUNCTIONS.register(name, hf_attention_forward)
    except AttributeError:
        ALL_ATTENTION_FUNCTIONS[name] = hf_attention_forward
    try:    # transformers >= 4.53 builds masks through a second registry; ours needs none (document ids drive the kernel)
        from transformers.masking_utils import ALL_MASK_ATTENTION_FUNCTIONS
        ALL_MASK_ATTENTION_FUNCTIONS.register(name, _no_mask)
    except Exception:
        pass
    return True


def _no_mask(*args, **kwargs):
    return None


# ---------------------------------------------------------------------------------------------------------------
# Liger-style patching of the stock HF Llama modules (SURVEY 8(b) "inner operator API")
# ---------------------------------------------------------------------------------------------------------------
class _HFRopeFn(torch.autograd.Function):
    """`apply_rotary_pos_emb(q [B,H,T,hd], k [B,KV,T,hd], cos [B,T,hd], sin)` (hf: modeling_llama.py:151-168) on the
    projection outputs as they lie in memory ([B,T,H,hd] behind the transpose), one kernel per tensor."""

    @staticmethod
    def forward(ctx, q, k, cos, sin):
        B, H, T, hd = q.shape
        KV = k.shape[1]
        half = hd // 2
        ct = cos[..., :half].reshape(B * T, half).to(torch.bfloat16).contiguous()
        st = sin[..., :half].reshape(B * T, half).to(torch.bfloat16).contiguous()
        q2 = q.transpose(1, 2).reshape(B * T, H * hd).to(torch.bfloat16, copy=True)
        k2 = k.transpose(1, 2).reshape(B * T, KV * hd).to(torch.bfloat16, copy=True)
        ops.rope_apply_(q2, ct, st, H, hd)
        ops.rope_apply_(k2, ct, st, KV, hd)
        ctx.save_for_backward(ct, st)
        ctx.dims = (B, T, H, KV, hd, q.dtype)
        return q2.view(B, T, H, hd).transpose(1, 2).to(q.dtype), k2.view(B, T, KV, hd).transpose(1, 2).to(k.dtype)

    @staticmethod
    def backward(ctx, dq, dk):
        ct, st = ctx.saved_tensors
        B, T, H, KV, hd, dt = ctx.dims
        dq2 = dq.transpose(1, 2).reshape(B * T, H * hd).to(torch.bfloat16, copy=True)
        dk2 = dk.transpose(1, 2).reshape(B * T, KV * hd).to(torch.bfloat16, copy=True)
        ops.rope_apply_(dq2, ct, st, H, hd, inverse=True)
        ops.rope_apply_(dk2, ct, st, KV, hd, inverse=True)
        return dq2.view(B, T, H, hd).transpose(1, 2).to(dt), dk2.view(B, T, KV, hd).transpose(1, 2).to(dt), None, None


def _hf_apply_rotary_pos_emb(q, k, cos, sin, position_ids=None, unsqueeze_dim=1):
    return _HFRopeFn.apply(q, k, cos, sin)


def _hf_rmsnorm_forward(self, hidden_states):
    return ops.rms_norm(hidden_states, self.weight, self.variance_epsilon)


def _hf_mlp_forward(self, x):
    if self.gate_proj.bias is not None or getattr(self.config, "hidden_act", "silu") != "silu":
        raise ops._lib.TouchNetB200Error("touchnet_b200 MLP kernel: SwiGLU (silu) without biases only")
    x2 = x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)
    h = ops.swiglu_mlp_in(x2, self.gate_proj.weight, self.up_proj.weight)
    return ops.linear(h, self.down_proj.weight).to(x.dtype)


def apply_b200_kernels_to_hf_llama(rope: bool = True, rms_norm: bool = True, swiglu: bool = True,
                                   attention: bool = True) -> None:
    """Swap the arithmetic of the STOCK transformers Llama modules for the B200 kernels, the way the reference lets
    Liger do it (`apply_liger_kernel_to_llama` in `additional_pre_init_fn`, ref: touchnet/models/llama/__init__.py:11-15):
    `LlamaRMSNorm.forward`, `LlamaMLP.forward`, `apply_rotary_pos_emb` are replaced in `transformers.models.llama.
    modeling_llama`, and "touchnet_b200" is registered as an attention implementation (select it with
    `config._attn_implementation = "touchnet_b200"` and wrap the call in `packed_document_ids(doc_ids)`).
    Module structure, parameters and state-dict keys stay HF's; head_dim must be 128 for the attention kernel."""
    from transformers.models.llama import modeling_llama
    if rope:
        modeling_llama.apply_rotary_pos_emb = _hf_apply_rotary_pos_emb
    if rms_norm:
        modeling_llama.LlamaRMSNorm.forward = _hf_rmsnorm_forward
    if swiglu:
        modeling_llama.LlamaMLP.forward = _hf_mlp_forward
    if attention:
        register_hf_attention()
