"""GPU parity: tcgen05 GEMM family through the C ABI vs fp32 torch matmul on the same bf16 inputs.
Tolerance: bf16 output rounding (2^-8 relative) + fp32 accumulation-order noise -> |err| <= 1e-2*max|ref| + 1e-2."""
import pytest
import torch

from tests.gpu_util import require_cuda
from touchnet_b200 import ops

pytestmark = pytest.mark.gpu

SHAPES = [(256, 256, 256), (256, 512, 128), (512, 256, 64), (128, 256, 64), (256, 128, 256), (512, 1024, 512), (200, 264, 1040), (1024, 4096, 400), (384, 1024, 4096)]


def _mk(M, N, K, dev):
    torch.manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    b = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    return a, b


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_forward_dgrad_wgrad(M, N, K):
    dev = require_cuda()
    a, b = _mk(M, N, K, dev)
    ref = a.float() @ b.float().t()
    tol = 1e-2 * float(ref.abs().max()) + 1e-2
    y = ops.gemm(a, b)                                           # forward  x·Wᵀ
    assert float((y.float() - ref).abs().max()) < tol
    dy = (torch.randn(M, N, device=dev) * 0.1).bfloat16()
    dx = ops.gemm(dy, b, b_mn=True)                              # dgrad    dy·W
    ref_dx = dy.float() @ b.float()
    assert float((dx.float() - ref_dx).abs().max()) < 1e-2 * float(ref_dx.abs().max()) + 1e-2
    dw = ops.gemm(dy, a, a_mn=True, b_mn=True, out_f32=True)     # wgrad    dyᵀ·x (fp32 straight from TMEM)
    ref_dw = dy.float().t() @ a.float()
    assert float((dw - ref_dw).abs().max()) < 1e-4 * float(ref_dw.abs().max()) + 1e-4
    dwb = ops.gemm(dy, a, a_mn=True, b_mn=True)                  # wgrad, bf16 (FSDP2 mixed-precision path)
    assert float((dwb.float() - ref_dw).abs().max()) < 1e-2 * float(ref_dw.abs().max()) + 1e-2


def test_gemm_residual_and_accumulate():
    dev = require_cuda()
    a, b = _mk(512, 1024, 512, dev)
    r = torch.randn(512, 1024, device=dev).bfloat16()
    y = ops.gemm(a, b, residual=r)
    ref = (a.float() @ b.float().t()).bfloat16().float() + r.float()
    assert float((y.float() - ref).abs().max()) <= 2 ** -6 * float(ref.abs().max())
    acc = torch.randn(1024, 512, device=dev)                      # fp32 gradient accumulation in place
    dy = (torch.randn(512, 1024, device=dev) * 0.1).bfloat16()
    want = acc + dy.float().t() @ a.float()
    ops.gemm(dy, a, a_mn=True, b_mn=True, out_f32=True, residual=acc, out=acc)
    assert float((acc - want).abs().max()) < 1e-3


@pytest.mark.parametrize("M,N,K", [(512, 1024, 512), (200, 264, 1040), (256, 14336, 256)])
def test_gemm_swiglu(M, N, K):
    dev = require_cuda()
    x, wg = _mk(M, N, K, dev)
    wu = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    g, u, h = ops.gemm_swiglu(x, wg, wu)
    g_ref = (x.float() @ wg.float().t())
    u_ref = (x.float() @ wu.float().t())
    assert float((g.float() - g_ref).abs().max()) < 1e-2 * float(g_ref.abs().max()) + 1e-2
    assert float((u.float() - u_ref).abs().max()) < 1e-2 * float(u_ref.abs().max()) + 1e-2
    # H must be exactly the unfused bf16 computation applied to the kernel's own (bf16) G and U
    h_ref = (torch.nn.functional.silu(g.float()).bfloat16().float() * u.float()).bfloat16()
    mism = (h != h_ref)
    assert float(mism.float().mean()) < 1e-3, "SwiGLU epilogue deviates from silu(g)*u on its own outputs"
    assert float((h.float() - h_ref.float()).abs().max()) <= 2 ** -6 * float(h_ref.float().abs().max()) + 1e-6


def test_gemm_rejects_bad_arguments():
    dev = require_cuda()
    from touchnet_b200._lib import TouchNetB200Error
    a = torch.zeros(128, 60, device=dev, dtype=torch.bfloat16)     # K=60: row stride not a multiple of 16 B
    b = torch.zeros(128, 60, device=dev, dtype=torch.bfloat16)
    with pytest.raises(TouchNetB200Error):
        ops.gemm(a, b)
    with pytest.raises(TouchNetB200Error):
        ops.gemm(torch.zeros(8, 8), torch.zeros(8, 8))             # CPU tensors: no CPU path


@pytest.mark.parametrize("M,d,nq,nkv", [(512, 512, 512, 256), (8192, 4096, 4096, 1024), (300, 1024, 256, 256)])
def test_fused_qkv_projection_forward_dgrad_wgrad(M, d, nq, nkv):
    """One-launch q/k/v projection with three separate weight tensors (segmented operands) == three separate GEMMs."""
    dev = require_cuda()
    torch.manual_seed(M + d)
    x = (torch.randn(M, d, device=dev) * 0.5).bfloat16()
    wq = (torch.randn(nq, d, device=dev) * 0.05).bfloat16()
    wk = (torch.randn(nkv, d, device=dev) * 0.05).bfloat16()
    wv = (torch.randn(nkv, d, device=dev) * 0.05).bfloat16()
    assert ops.qkv_fusable(M, nq, nkv)
    qkv = ops.gemm_qkv_fwd(x, wq, wk, wv)
    ref = torch.cat([ops.gemm(x, wq), ops.gemm(x, wk), ops.gemm(x, wv)], dim=1)
    assert torch.equal(qkv, ref)                                   # same tiles, same k order: bit identical
    dqkv = (torch.randn(M, nq + 2 * nkv, device=dev) * 0.1).bfloat16()
    dx = ops.gemm_qkv_dgrad(dqkv, wq, wk, wv)
    dx_ref = dqkv.float() @ torch.cat([wq, wk, wv]).float()
    assert float((dx.float() - dx_ref).abs().max()) < 1e-2 * float(dx_ref.abs().max()) + 1e-2
    dws = ops.gemm_qkv_wgrad(dqkv, x, True, nq, nkv)
    offs = [0, nq, nq + nkv, nq + 2 * nkv]
    for i, dw in enumerate(dws):
        want = dqkv[:, offs[i]:offs[i + 1]].float().t() @ x.float()
        assert dw.dtype == torch.float32 and float((dw - want).abs().max()) < 1e-4 * float(want.abs().max()) + 1e-4
    dws_b = ops.gemm_qkv_wgrad(dqkv, x, False, nq, nkv)              # bf16 parameters (FSDP2 mixed precision)
    assert all(t.dtype == torch.bfloat16 for t in dws_b)


@pytest.mark.parametrize("M,d,H,KV", [(512, 512, 4, 2), (8192, 4096, 32, 8), (300, 1024, 2, 2)])
def test_fused_qkv_rope_epilogue_is_bit_identical_to_unfused(M, d, H, KV):
    """RoPE in the QKV GEMM epilogue == GEMM followed by the in-place RoPE kernel (same bf16 rounding points)."""
    dev = require_cuda()
    from oracle import model_oracle as mo
    torch.manual_seed(M)
    nq, nkv = H * 128, KV * 128
    x = (torch.randn(M, d, device=dev) * 0.5).bfloat16()
    wq = (torch.randn(nq, d, device=dev) * 0.05).bfloat16()
    wk = (torch.randn(nkv, d, device=dev) * 0.05).bfloat16()
    wv = (torch.randn(nkv, d, device=dev) * 0.05).bfloat16()
    pos = (torch.arange(M, device=dev) % 700)[None]
    inv, sc = mo.rope_inv_freq(mo.OracleConfig(d, 8, 1, H, KV, 128, 8, rope_theta=500000.0))
    cos, sin = ops.rope_table(pos, inv.to(dev), sc)
    fused = ops.gemm_qkv_fwd(x, wq, wk, wv, rope=(cos, sin))
    plain = ops.gemm_qkv_fwd(x, wq, wk, wv)
    ops.rope_apply_(plain[:, :nq], cos, sin, H, 128)
    ops.rope_apply_(plain[:, nq:nq + nkv], cos, sin, KV, 128)
    assert torch.equal(fused, plain)


@pytest.mark.parametrize("M,N,K", [(512, 1024, 256), (384, 520, 192), (8192, 14336, 4096)])
def test_gemm_dswiglu_equals_dgrad_then_swiglu_backward(M, N, K):
    """Down-proj dgrad with the SwiGLU backward in its epilogue == tn_gemm_bf16 (dH, bf16) followed by tn_swiglu_bwd_bf16,
    bit for bit (same rounding points), incl. ragged edges and the Llama-3-8B shape (8192, 14336, 4096)."""
    dev = require_cuda()
    torch.manual_seed(M + N + K)
    dy = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    wd = (torch.randn(K, N, device=dev) * 0.05).bfloat16()          # down_proj.weight [d, ffn]
    g = torch.randn(M, N, device=dev).bfloat16()
    u = torch.randn(M, N, device=dev).bfloat16()
    dg_ref, du_ref = ops.swiglu_bwd(g, u, ops.gemm(dy, wd, b_mn=True))
    ops._FUSE_DSWIGLU, saved = True, ops._FUSE_DSWIGLU            # the fused launch is opt-in (TN_FUSED_DSWIGLU=1)
    try:
        dg, du = ops.gemm_dswiglu(dy, wd, g, u)
    finally:
        ops._FUSE_DSWIGLU = saved
    assert torch.equal(dg, dg_ref) and torch.equal(du, du_ref)
    # and against fp32 math: dG = dH * u * silu'(g), dU = dH * silu(g)
    dh = dy.float() @ wd.float()
    sig = torch.sigmoid(g.float())
    du32 = dh * (g.float() * sig)
    dg32 = dh * u.float() * (sig * (1 + g.float() * (1 - sig)))
    assert float((du.float() - du32).abs().max()) <= 2e-2 * float(du32.abs().max()) + 1e-2
    assert float((dg.float() - dg32).abs().max()) <= 2e-2 * float(dg32.abs().max()) + 1e-2
