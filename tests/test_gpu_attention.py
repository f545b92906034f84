"""GPU parity: packed-document attention (forward + backward) through the C ABI vs the oracle's dense-mask attention
(oracle/model_oracle.py::attention, pinned to HF eager / FlexAttention semantics).
Tolerance (bf16 kernel vs fp32 oracle on the same bf16 inputs): O and dQ/dK/dV relative L2 error < 2e-2 and max error
< 3e-2*max|ref| + 3e-2; LSE abs error < 2e-3; padding rows exactly 0."""
import math

import pytest
import torch

from oracle import model_oracle as mo
from tests.gpu_util import max_err, packed_doc_ids, rel_err, require_cuda
from touchnet_b200 import ops

pytestmark = pytest.mark.gpu

CASES = {
    # name: (B, T, H, KV, lens_per_row)
    "multi_doc_pad": (2, 512, 4, 2, [[100, 200, 150], [300, 50]]),
    "single_doc_full_blocks": (1, 512, 2, 2, [[512]]),
    "ragged_T": (2, 300, 2, 1, [[120, 180], [33, 90, 100]]),
    "gqa4": (1, 384, 8, 2, [[200, 100, 84]]),
    "long_doc": (1, 2048, 2, 1, [[1500, 500]]),
    "tiny_docs": (1, 256, 2, 2, [[3, 1, 7, 2, 60, 1, 1, 50, 100]]),
    "all_pad_row": (2, 256, 2, 1, [[256], []]),
}


def _run(name, custom_doc=None):
    dev = require_cuda()
    B, T, H, KV, lens = CASES[name]
    torch.manual_seed(hash(name) % 1000)
    doc, _ = packed_doc_ids(B, T, lens, dev)
    if custom_doc is not None:
        doc = custom_doc.to(dev)
    q = torch.randn(B * T, H * 128, device=dev).bfloat16()
    k = torch.randn(B * T, KV * 128, device=dev).bfloat16()
    v = torch.randn(B * T, KV * 128, device=dev).bfloat16()
    scale = 1 / math.sqrt(128)
    plan = ops.AttnPlan(doc)
    o, lse = ops.attn_fwd(q, k, v, plan, H, KV, scale)
    torch.cuda.synchronize()
    # oracle
    qf = q.float().view(B, T, H, 128).transpose(1, 2).detach().requires_grad_(True)
    kf = k.float().view(B, T, KV, 128).transpose(1, 2).detach().requires_grad_(True)
    vf = v.float().view(B, T, KV, 128).transpose(1, 2).detach().requires_grad_(True)
    allow = mo.doc_causal_allow(doc)
    o_ref, lse_ref = mo.attention(qf, kf, vf, allow, scale)
    o_ref2 = o_ref.reshape(B * T, H * 128)
    valid = (doc > 0).reshape(-1)
    assert torch.all(o[~valid] == 0), "padding query rows must be exactly zero (FlexAttention semantics)"
    assert rel_err(o[valid].float(), o_ref2[valid]) < 2e-2, (name, rel_err(o[valid].float(), o_ref2[valid]))
    assert max_err(o.float(), o_ref2) < 3e-2 * float(o_ref2.abs().max()) + 3e-2
    vm = (doc > 0)[:, None, :].expand(B, H, T)
    assert max_err(lse[vm], lse_ref[vm]) < 2e-3, max_err(lse[vm], lse_ref[vm])
    assert torch.isinf(lse[~vm]).all()
    # backward
    do = torch.randn(B * T, H * 128, device=dev).bfloat16()
    dq, dk, dv = ops.attn_bwd(q, k, v, o, do, lse, plan, H, KV, scale)
    torch.cuda.synchronize()
    o_ref.backward(do.float().view(B, T, H, 128))
    dq_ref = qf.grad.transpose(1, 2).reshape(B * T, H * 128)
    dk_ref = kf.grad.transpose(1, 2).reshape(B * T, KV * 128)
    dv_ref = vf.grad.transpose(1, 2).reshape(B * T, KV * 128)
    for nm, a, r in (("dq", dq, dq_ref), ("dk", dk, dk_ref), ("dv", dv, dv_ref)):
        assert torch.isfinite(a.float()).all(), (name, nm)
        assert rel_err(a.float(), r) < 2e-2, (name, nm, rel_err(a.float(), r))
        assert max_err(a.float(), r) < 3e-2 * float(r.abs().max()) + 3e-2, (name, nm)
    assert torch.all(dq[~valid] == 0)


@pytest.mark.parametrize("name", list(CASES))
def test_attention_parity(name):
    _run(name)


def test_attention_noncanonical_doc_ids():
    """Ids that repeat non-contiguously (1,2,1,...) are legal for the reference's mask_mod; the kernel must fall back to
    exact element-wise doc-id masking over the whole causal range."""
    B, T = 1, 384
    doc = torch.ones(B, T, dtype=torch.int64)
    doc[0, 100:200] = 2
    doc[0, 300:] = 0
    doc[0, 350:] = 3   # a document after padding
    CASES["noncanon"] = (B, T, 2, 1, [[T]])
    _run("noncanon", custom_doc=doc)


def test_attention_meta_ranges():
    dev = require_cuda()
    doc, _ = packed_doc_ids(1, 1024, [[300, 500, 100]], dev)       # docs [0,300) [300,800) [800,900), pad [900,1024)
    plan = ops.AttnPlan(doc)
    torch.cuda.synchronize()
    meta = plan.meta[: 8 * 4].view(8, 4).cpu()
    assert int(plan.meta[8 * 4]) == 1                              # canonical
    # q block 2 (rows 256..383) starts in doc 1 (start 0) -> kv_lo = 0; q block 3 (384..511) is doc 2 (start 300) -> 2
    assert meta[2].tolist()[:2] == [0, 3] and meta[3].tolist()[:2] == [2, 4]
    assert meta[6].tolist()[:2] == [2, 7]                          # rows 768.. start in doc 2 (start 300)
    assert meta[7].tolist()[:2] == [6, 8]                          # rows 896..899 are doc 3 (start 800)
    # kv block 2 (cols 256..383): last valid col 383 is doc 2, run ends at 800 -> q blocks up to 6 (exclusive 7)
    assert meta[2].tolist()[2] == 7 and meta[0].tolist()[2] == 3 and meta[7].tolist()[2] == 8


def test_attention_autograd_function():
    dev = require_cuda()
    B, T, H, KV = 1, 256, 2, 1
    doc, pos = packed_doc_ids(B, T, [[100, 156]], dev)
    inv, sc = mo.rope_inv_freq(mo.OracleConfig(256, 8, 1, H, KV, 128, 8, rope_theta=10000.0))
    cos, sin = ops.rope_table(pos, inv.to(dev), sc)
    q = torch.randn(B * T, H * 128, device=dev).bfloat16().requires_grad_(True)
    k = torch.randn(B * T, KV * 128, device=dev).bfloat16().requires_grad_(True)
    v = torch.randn(B * T, KV * 128, device=dev).bfloat16().requires_grad_(True)
    q_in, k_in = q.detach().clone(), k.detach().clone()
    o = ops.PackedAttentionFn.apply(q.clone(), k.clone(), v, cos, sin, ops.AttnPlan(doc), H, KV, 1 / math.sqrt(128))
    o.float().square().sum().backward()
    # oracle: rope + attention, fp32
    qf = q_in.float().view(B, T, H, 128).transpose(1, 2).requires_grad_(True)
    kf = k_in.float().view(B, T, KV, 128).transpose(1, 2).requires_grad_(True)
    vf = v.detach().float().view(B, T, KV, 128).transpose(1, 2).requires_grad_(True)
    cr, sr = mo.rope_cos_sin(pos, inv.to(dev), sc, torch.float32)
    qr, kr = mo.apply_rope(qf, kf, cr, sr)
    o_ref, _ = mo.attention(qr, kr, vf, mo.doc_causal_allow(doc), 1 / math.sqrt(128))
    o_ref.square().sum().backward()
    assert rel_err(o.float(), o_ref.reshape(B * T, -1)) < 2e-2
    assert rel_err(q.grad.float(), qf.grad.transpose(1, 2).reshape(B * T, -1)) < 3e-2
    assert rel_err(k.grad.float(), kf.grad.transpose(1, 2).reshape(B * T, -1)) < 3e-2
    assert rel_err(v.grad.float(), vf.grad.transpose(1, 2).reshape(B * T, -1)) < 3e-2


def test_attention_matches_torch_flex_attention_with_hf_block_mask():
    """The kernel the reference actually runs on a GPU: torch.nn.attention.flex_attention driven by HF's
    make_flex_block_causal_mask(document ids) (hf: integrations/flex_attention.py:136-247, :262-364).  Compared at the op
    level on identical bf16 q/k/v; bf16-vs-bf16 tolerance: rel L2 < 1e-2."""
    dev = require_cuda()
    from torch.nn.attention.flex_attention import create_block_mask, flex_attention
    B, T, H, KV = 2, 512, 4, 2
    doc, _ = packed_doc_ids(B, T, [[100, 200, 150], [300, 50]], dev)
    torch.manual_seed(5)
    q = torch.randn(B, T, H, 128, device=dev).bfloat16()
    k = torch.randn(B, T, KV, 128, device=dev).bfloat16()
    v = torch.randn(B, T, KV, 128, device=dev).bfloat16()
    scale = 1 / math.sqrt(128)
    try:
        from transformers.integrations.flex_attention import make_flex_block_causal_mask
        bm = make_flex_block_causal_mask(doc)
    except Exception:
        def mask_mod(b, h, qi, ki):   # restated from hf: integrations/flex_attention.py (causal & same doc & q not pad)
            return (qi >= ki) & (doc[b, qi] == doc[b, ki]) & (doc[b, qi] > 0)
        bm = create_block_mask(mask_mod, B, None, T, T, device=dev)
    o_ref = flex_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), block_mask=bm, enable_gqa=True,
                           scale=scale).transpose(1, 2).reshape(B * T, H * 128)
    plan = ops.AttnPlan(doc)
    o, _ = ops.attn_fwd(q.view(B * T, -1), k.view(B * T, -1), v.view(B * T, -1), plan, H, KV, scale)
    valid = (doc > 0).reshape(-1)
    assert rel_err(o[valid].float(), o_ref[valid].float()) < 1e-2
    assert torch.all(o[~valid] == 0) and torch.all(o_ref[~valid] == 0)      # both give exact zeros on padding rows


@pytest.mark.parametrize("T,H,KV,lens,cp", [(1024, 4, 2, [[300, 500, 100]], 2), (1024, 2, 1, [[1024]], 4),
                                             (768, 4, 4, [[100, 50, 400, 218]], 3)])
def test_context_parallel_query_windows(T, H, KV, lens, cp):
    """Context parallelism shards the sequence over `cp` ranks (ref: touchnet/utils/distributed.py:292-315,
    touchnet/bin/train.py:363-387).  Kernel-level contract, checked on one GPU: running the kernels on each rank's query
    window (K/V global, as after the K/V all-gather) reproduces the rows of the unsharded result, and the per-window
    dK/dV partial sums add up to the unsharded dK/dV (what the reduce-scatter computes)."""
    dev = require_cuda()
    B = len(lens)
    torch.manual_seed(T + cp)
    doc, _ = packed_doc_ids(B, T, lens, dev)
    q = torch.randn(B, T, H * 128, device=dev).bfloat16()
    k = torch.randn(B * T, KV * 128, device=dev).bfloat16()
    v = torch.randn(B * T, KV * 128, device=dev).bfloat16()
    do = torch.randn(B, T, H * 128, device=dev).bfloat16()
    scale = 1 / math.sqrt(128)
    full = ops.AttnPlan(doc)
    o_full, lse_full = ops.attn_fwd(q.view(B * T, -1), k, v, full, H, KV, scale)
    dq_full, dk_full, dv_full = ops.attn_bwd(q.view(B * T, -1), k, v, o_full, do.view(B * T, -1), lse_full, full, H, KV, scale)
    Tq = T // cp
    assert Tq % 128 == 0
    dk_sum = torch.zeros_like(dk_full, dtype=torch.float32)
    dv_sum = torch.zeros_like(dv_full, dtype=torch.float32)
    for r in range(cp):
        win = ops.AttnPlan(doc, Tq=Tq, q_blk_off=r * Tq // 128)
        ql = q[:, r * Tq:(r + 1) * Tq].reshape(B * Tq, -1).contiguous()
        dol = do[:, r * Tq:(r + 1) * Tq].reshape(B * Tq, -1).contiguous()
        o_l, lse_l = ops.attn_fwd(ql, k, v, win, H, KV, scale)
        ref_rows = o_full.view(B, T, -1)[:, r * Tq:(r + 1) * Tq].reshape(B * Tq, -1)
        assert torch.equal(o_l, ref_rows)                                   # same tiles, same order: bit identical
        assert torch.equal(lse_l, lse_full[:, :, r * Tq:(r + 1) * Tq])
        dq_l, dk_l, dv_l = ops.attn_bwd(ql, k, v, o_l, dol, lse_l, win, H, KV, scale)
        assert torch.equal(dq_l, dq_full.view(B, T, -1)[:, r * Tq:(r + 1) * Tq].reshape(B * Tq, -1))
        dk_sum += dk_l.float()
        dv_sum += dv_l.float()
    assert rel_err(dk_sum, dk_full.float()) < 1e-2 and rel_err(dv_sum, dv_full.float()) < 1e-2


def test_backward_fused_inverse_rope_matches_separate_kernel():
    """dQ/dK with the inverse RoPE applied in the attention-backward epilogue (fp32, one rounding) vs the separate
    in-place kernel on the bf16 gradients (several roundings): equal within bf16 noise; dV untouched."""
    dev = require_cuda()
    B, T, H, KV = 2, 384, 4, 2
    doc, pos = packed_doc_ids(B, T, [[100, 200, 50], [384]], dev)
    inv, sc = mo.rope_inv_freq(mo.OracleConfig(512, 8, 1, H, KV, 128, 8, rope_theta=10000.0))
    cos, sin = ops.rope_table(pos, inv.to(dev), sc)
    torch.manual_seed(9)
    q = torch.randn(B * T, H * 128, device=dev).bfloat16()
    k = torch.randn(B * T, KV * 128, device=dev).bfloat16()
    v = torch.randn(B * T, KV * 128, device=dev).bfloat16()
    do = torch.randn(B * T, H * 128, device=dev).bfloat16()
    plan = ops.AttnPlan(doc)
    scale = 1 / math.sqrt(128)
    o, lse = ops.attn_fwd(q, k, v, plan, H, KV, scale)
    dq, dk, dv = ops.attn_bwd(q, k, v, o, do, lse, plan, H, KV, scale)
    ops.rope_apply_(dq, cos, sin, H, 128, inverse=True)
    ops.rope_apply_(dk, cos, sin, KV, 128, inverse=True)
    dq_f, dk_f, dv_f = ops.attn_bwd(q, k, v, o, do, lse, plan, H, KV, scale, rope=(cos, sin))
    assert torch.equal(dv_f, dv)
    assert rel_err(dq_f.float(), dq.float()) < 1e-2 and rel_err(dk_f.float(), dk.float()) < 1e-2
