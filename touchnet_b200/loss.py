"""Pack-loss cross-entropy and accuracy behind the reference's `loss_fn` / `acc_fn` slots of TrainSpec
(ref: touchnet/utils/train_spec.py:37-38).

    cross_entropy_loss(pred, labels, sentence_lens, num_sentence) -> (loss_per_sample, loss_per_token)
        ref: touchnet/loss/cross_entropy.py:12-50 (+ the compiled fp32-upcast CE of touchnet/loss/__init__.py:7-28)
    accuracy(pred, labels) -> fraction of non-ignored positions whose argmax equals the label
        ref: touchnet/utils/metrics.py:26-50

The CUDA kernels (csrc/loss.cu) stream the bf16 logits once forward (online logsumexp + argmax) and once backward,
overwriting them in place with the logit gradient; no fp32 copy of the [B,T,V] tensor is ever made.  In-place is safe in
the reference's loop order: loss_fn and acc_fn run, `del pred`, then backward (ref: touchnet/bin/train.py:447-455).
"""
from __future__ import annotations

import weakref

import torch

from . import _lib

# argmax of the logits the loss kernel saw last: (weak reference to the caller's logits tensor, its version, argmax).
# Keyed on the live tensor OBJECT, not on its address - a freed tensor's address is reused by the next logits.
_LAST_ARGMAX = None


def _check_logits(logits: torch.Tensor) -> None:
    if not logits.is_cuda or logits.dtype != torch.bfloat16:
        raise _lib.TouchNetB200Error("pack-loss CE needs bf16 CUDA logits (no CPU path)")


def _st() -> int:
    return torch.cuda.current_stream().cuda_stream


class _PackCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, sentence_lens, inv_num_sentence):
        _check_logits(logits)
        V = logits.shape[-1]
        x = logits.view(-1, V)
        assert x.stride(1) == 1
        M = x.shape[0]
        lab = labels.reshape(-1).contiguous()
        sl = sentence_lens.reshape(-1).contiguous()
        lse = torch.empty(M, dtype=torch.float32, device=x.device)
        ce = torch.empty(M, dtype=torch.float32, device=x.device)
        am = torch.empty(M, dtype=torch.int32, device=x.device)
        _lib.call("tn_pack_ce_fwd_bf16", x.data_ptr(), x.stride(0), lab.data_ptr(), lse.data_ptr(), ce.data_ptr(),
                  am.data_ptr(), M, V, _st())
        ctx.save_for_backward(lab, sl, lse)
        ctx.logits = x                      # NOT save_for_backward: backward overwrites it in place on purpose
        ctx.inv_ns = inv_num_sentence
        ctx.shape = logits.shape
        ctx.mark_non_differentiable(ce, am)
        loss_per_sample = (ce / sl.float()).sum() * inv_num_sentence
        return loss_per_sample, ce, am

    @staticmethod
    def backward(ctx, g_loss, _g_ce, _g_am):
        lab, sl, lse = ctx.saved_tensors
        x = ctx.logits
        g = g_loss.reshape(1).float().contiguous()
        _lib.call("tn_pack_ce_bwd_bf16", x.data_ptr(), x.stride(0), lab.data_ptr(), sl.data_ptr(), lse.data_ptr(),
                  g.data_ptr(), float(ctx.inv_ns), x.shape[0], x.shape[1], _st())
        torch.autograd.graph.increment_version(x)   # the kernel overwrote the caller's logits with their gradient: make
        ctx.logits = None                           # autograd (and accuracy()'s cache key) see the mutation
        return x.view(ctx.shape), None, None, None


def cross_entropy_loss(pred: torch.Tensor, labels: torch.Tensor, sentence_lens: torch.Tensor, num_sentence: int,
                       ignore_index: int = -100):
    """Same contract as ref: touchnet/loss/cross_entropy.py:12-50.  Returns (loss_per_sample, loss_per_token)."""
    assert ignore_index < 0, "labels outside [0, V) are ignored (the reference uses -100)"
    B = pred.shape[0]
    global _LAST_ARGMAX
    loss_per_sample, ce, am = _PackCEFn.apply(pred, labels, sentence_lens, 1.0 / max(int(num_sentence), 1))
    _LAST_ARGMAX = (weakref.ref(pred), pred._version, am)
    with torch.no_grad():
        num_tokens = (labels != ignore_index).sum()
        tot = ce.sum()
        loss_per_token = torch.where((tot > 1e-6) & (num_tokens > 0), tot / num_tokens.clamp(min=1), torch.zeros_like(tot))
    return loss_per_sample, loss_per_token


def accuracy(pred: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
    """Same contract as ref: touchnet/utils/metrics.py:26-50; reuses the argmax the loss kernel already produced."""
    V = pred.shape[-1]
    x = pred.view(-1, V)
    ent = _LAST_ARGMAX
    if ent is not None and ent[0]() is pred and ent[1] == pred._version and ent[2].numel() == x.shape[0]:
        am = ent[2]
    else:
        _check_logits(pred)
        M = x.shape[0]
        lse = torch.empty(M, dtype=torch.float32, device=x.device)
        ce = torch.empty_like(lse)
        am = torch.empty(M, dtype=torch.int32, device=x.device)
        lab = labels.reshape(-1).contiguous()
        _lib.call("tn_pack_ce_fwd_bf16", x.data_ptr(), x.stride(0), lab.data_ptr(), lse.data_ptr(), ce.data_ptr(),
                  am.data_ptr(), M, V, _st())
    lab = labels.reshape(-1)
    mask = lab != ignore_index
    num = ((am.long() == lab) & mask).sum()
    den = mask.sum()
    return torch.where(den > 0, num / den.clamp(min=1), torch.zeros_like(num, dtype=torch.float32)).detach()
