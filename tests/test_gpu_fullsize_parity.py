"""Full-size VALUE parity (BASELINE.json configurations), through the C ABI:

  * attention forward + backward at cfg 2 (Llama-3-8B, T=8192, H=32, KV=8, the bench's ASR document mix), cfg 3
    (Qwen2-Audio-7B shape: MHA, B=2, T=4096) and cfg 4 (one 32768-token document, H=32, KV=8) against
      (a) the fp32 oracle evaluated on the GPU one head / one block of query rows at a time
          (oracle/model_oracle.py::attention_chunked - the dense-mask restatement, pinned to `attention` on CPU), and
      (b) COMPILED torch flex_attention driven by HF's make_flex_block_causal_mask - the kernel the reference runs
          (touchnet/bin/train.py:129-131, hf: integrations/flex_attention.py:136-247, :262-364);
  * the tcgen05 GEMMs at the shapes that carry 80 % of the step - gate/up (8192, 14336, 4096), down (8192, 4096, 14336),
    lm_head (8192, 128256, 4096) - forward, dgrad and wgrad against fp32 matmuls of the same bf16 operands.

Tolerances (stated per comparison below): bf16 kernel vs fp32 oracle on identical bf16 inputs - relative L2 < 2e-2, LSE
abs < 2e-3, padding rows exactly 0; vs compiled flex (bf16 vs bf16) relative L2 < 1e-2; GEMM |err| <= 1e-2*max|ref|."""
import math

import pytest
import torch

from oracle import model_oracle as mo
from tests.gpu_util import max_err, packed_doc_ids, rel_err, require_cuda
from touchnet_b200 import batching, ops

pytestmark = pytest.mark.gpu
SCALE = 1 / math.sqrt(128)


def _asr_doc_ids(B, T, seed, dev):
    """The bench workload's packing (bench.py::make_host_batch): audio+text documents, utterances U[1,30] s."""
    buf, _ = batching.plan_audio_text_batch(seed, B, T, 128256, stride=4, max_s=30.0)
    return buf["attention_mask"].to(dev)


def _cases():
    return {
        "cfg2_llama3_8b_T8192_asr_docs": (1, 8192, 32, 8, "asr"),
        "cfg3_qwen2_audio_T4096_B2_mha": (2, 4096, 32, 32, "asr"),
        "cfg4_T32768_one_document": (1, 32768, 32, 8, "one"),
    }


def _inputs(name):
    dev = require_cuda()
    B, T, H, KV, kind = _cases()[name]
    doc = _asr_doc_ids(B, T, 2025, dev) if kind == "asr" else packed_doc_ids(B, T, [[T]], dev)[0]
    g = torch.Generator(device="cpu").manual_seed(len(name))
    mk = lambda c: torch.randn(B * T, c * 128, generator=g).to(dev).bfloat16()
    return dev, B, T, H, KV, doc, mk(H), mk(KV), mk(KV), mk(H)


def _ours(B, T, H, KV, doc, q, k, v, do):
    plan = ops.AttnPlan(doc)
    o, lse = ops.attn_fwd(q, k, v, plan, H, KV, SCALE)
    dq, dk, dv = ops.attn_bwd(q, k, v, o, do, lse, plan, H, KV, SCALE)
    torch.cuda.synchronize()
    return o, lse, dq, dk, dv


@pytest.mark.parametrize("name", list(_cases()))
def test_attention_full_size_vs_chunked_fp32_oracle(name):
    dev, B, T, H, KV, doc, q, k, v, do = _inputs(name)
    o, lse, dq, dk, dv = _ours(B, T, H, KV, doc, q, k, v, do)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False                 # the oracle is fp32, not tf32
    try:
        o_r, lse_r, dq_r, dk_r, dv_r = mo.attention_chunked(q, k, v, doc, H, KV, SCALE, do, q_chunk=4096)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    valid = (doc > 0).reshape(-1)
    assert torch.all(o[~valid] == 0) and torch.all(dq[~valid] == 0), "padding query rows must be exactly zero"
    vm = (doc > 0)[:, None, :].expand(B, H, T)
    assert torch.isinf(lse[~vm]).all() and torch.isinf(lse_r[~vm]).all()
    assert max_err(lse[vm], lse_r[vm]) < 2e-3, (name, "lse", max_err(lse[vm], lse_r[vm]))
    for nm, a, r in (("o", o, o_r), ("dq", dq, dq_r), ("dk", dk, dk_r), ("dv", dv, dv_r)):
        assert torch.isfinite(a.float()).all(), (name, nm)
        e = rel_err(a.float(), r)
        assert e < 2e-2, (name, nm, e)
        assert max_err(a.float(), r) < 3e-2 * float(r.abs().max()) + 3e-2, (name, nm)


@pytest.mark.parametrize("name", list(_cases()))
def test_attention_full_size_vs_compiled_flex_attention(name):
    """bf16 vs bf16 against the compiled FlexAttention kernel with HF's BlockMask; skipped only if torch.compile cannot
    build the kernel in this image."""
    dev, B, T, H, KV, doc, q, k, v, do = _inputs(name)
    try:
        from torch.nn.attention.flex_attention import flex_attention
        from transformers.integrations.flex_attention import make_flex_block_causal_mask
        bm = make_flex_block_causal_mask(doc)
        flex = torch.compile(flex_attention, dynamic=False)
        q4 = q.view(B, T, H, 128).transpose(1, 2).detach().requires_grad_(True)
        k4 = k.view(B, T, KV, 128).transpose(1, 2).detach().requires_grad_(True)
        v4 = v.view(B, T, KV, 128).transpose(1, 2).detach().requires_grad_(True)
        o_f = flex(q4, k4, v4, block_mask=bm, enable_gqa=(H != KV), scale=SCALE)
        o_f.backward(do.view(B, T, H, 128).transpose(1, 2))
        torch.cuda.synchronize()
    except Exception as e:  # pragma: no cover - depends on the box's inductor/triton toolchain
        pytest.skip(f"compiled flex_attention not runnable here: {type(e).__name__}: {str(e)[:300]}")
    o, lse, dq, dk, dv = _ours(B, T, H, KV, doc, q, k, v, do)
    valid = (doc > 0).reshape(-1)
    tok = lambda x, heads: x.detach().transpose(1, 2).reshape(B * T, heads * 128)
    o_r, dq_r, dk_r, dv_r = tok(o_f, H), tok(q4.grad, H), tok(k4.grad, KV), tok(v4.grad, KV)
    assert torch.all(o[~valid] == 0) and torch.all(o_r[~valid] == 0)          # both: exact zeros on padding rows
    for nm, a, r in (("o", o, o_r), ("dq", dq, dq_r), ("dk", dk, dk_r), ("dv", dv, dv_r)):
        e = rel_err(a[valid].float(), r[valid].float()) if nm in ("o", "dq") else rel_err(a.float(), r.float())
        assert e < 1e-2, (name, nm, e)


# ---------------------------------------------------------------------------------------------------------------
# GEMMs at the Llama-3-8B shapes (M = B*T = 8192 tokens)
# ---------------------------------------------------------------------------------------------------------------
GEMM_SHAPES = {"gate_up": (8192, 14336, 4096), "down": (8192, 4096, 14336), "lm_head": (8192, 128256, 4096)}


@pytest.mark.parametrize("name", list(GEMM_SHAPES))
def test_gemm_full_size_forward_dgrad_wgrad(name):
    dev = require_cuda()
    M, N, K = GEMM_SHAPES[name]
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = (torch.randn(M, K, generator=g) * 0.5).to(dev).bfloat16()        # activations
    b = (torch.randn(N, K, generator=g) * 0.02).to(dev).bfloat16()       # nn.Linear.weight [out, in]
    dy = (torch.randn(M, N, generator=g) * 0.1).to(dev).bfloat16()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        # forward  y = a . b^T
        y = ops.gemm(a, b)
        ref = a.float() @ b.float().t()
        assert max_err(y.float(), ref) <= 1e-2 * float(ref.abs().max()) + 1e-2, (name, "fwd")
        assert rel_err(y.float(), ref) < 4e-3, (name, "fwd", rel_err(y.float(), ref))
        del y, ref
        # dgrad    dx = dy . b
        dx = ops.gemm(dy, b, b_mn=True)
        ref = dy.float() @ b.float()
        assert max_err(dx.float(), ref) <= 1e-2 * float(ref.abs().max()) + 1e-2, (name, "dgrad")
        assert rel_err(dx.float(), ref) < 4e-3, (name, "dgrad")
        del dx, ref
        # wgrad    dW = dy^T . a, fp32 straight from the TMEM accumulator (fp32 master weights)
        dw = ops.gemm(dy, a, a_mn=True, b_mn=True, out_f32=True)
        ref = dy.float().t() @ a.float()
        assert max_err(dw, ref) <= 2e-4 * float(ref.abs().max()) + 1e-4, (name, "wgrad")
        assert rel_err(dw, ref) < 1e-4, (name, "wgrad", rel_err(dw, ref))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def test_gemm_swiglu_full_size():
    """gate+up GEMM with the SwiGLU epilogue at (8192, 14336, 4096): G, U vs fp32, H exactly silu(G)*U of its own G, U."""
    dev = require_cuda()
    M, N, K = GEMM_SHAPES["gate_up"]
    g = torch.Generator(device="cpu").manual_seed(3)
    x = (torch.randn(M, K, generator=g) * 0.5).to(dev).bfloat16()
    wg = (torch.randn(N, K, generator=g) * 0.02).to(dev).bfloat16()
    wu = (torch.randn(N, K, generator=g) * 0.02).to(dev).bfloat16()
    G, U, Hh = ops.gemm_swiglu(x, wg, wu)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for nm, got, w in (("G", G, wg), ("U", U, wu)):
            ref = x.float() @ w.float().t()
            assert max_err(got.float(), ref) <= 1e-2 * float(ref.abs().max()) + 1e-2, nm
            del ref
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    h_ref = (torch.nn.functional.silu(G.float()).bfloat16().float() * U.float()).bfloat16()
    assert float((Hh != h_ref).float().mean()) < 1e-3
