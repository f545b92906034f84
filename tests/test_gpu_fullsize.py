"""Full-size (BASELINE.json configurations) checks of the attention kernels through size-independent properties - the
dense-mask oracle cannot run at these sizes (T^2 scores), so what is verified is what the domain guarantees:

  * row-sum: with V == 1 every valid output row is exactly 1 (softmax rows sum to one), padding rows exactly 0;
  * document independence: a document's outputs / gradients inside a packed row equal the same document run alone
    (different block partition, so bf16-level tolerance, not bit equality);
  * causality: changing K/V of later positions leaves earlier outputs BIT-identical;
  * linearity of the backward in dO (exact for a power-of-two factor);
  * run-to-run determinism (bitwise) of forward and backward - no atomics anywhere on the path.

Sizes: cfg 2 (T=8192, H=32, KV=8, ASR-like document lengths), cfg 3 (MHA, T=4096, B=2), cfg 4 (T=32768, one document)."""
import math

import pytest
import torch

from tests.gpu_util import packed_doc_ids, rel_err, require_cuda
from touchnet_b200 import ops

pytestmark = pytest.mark.gpu
SCALE = 1 / math.sqrt(128)


def _doc_lens(T, seed, lo=40, hi=900, pad=300):
    g = torch.Generator().manual_seed(seed)
    lens, left = [], T - pad
    while left > lo:
        n = int(torch.randint(lo, hi, (1,), generator=g))
        n = min(n, left)
        lens.append(n)
        left -= n
    return lens


def _qkv(B, T, H, KV, dev, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    mk = lambda c: (torch.randn(B * T, c * 128, generator=g) * 1.0).to(dev).bfloat16()
    return mk(H), mk(KV), mk(KV)


@pytest.mark.parametrize("B,T,H,KV", [(1, 8192, 32, 8), (2, 4096, 32, 32)])
def test_packed_row_properties_at_full_size(B, T, H, KV):
    dev = require_cuda()
    lens = [_doc_lens(T, 7 + b) for b in range(B)]
    doc, _ = packed_doc_ids(B, T, lens, dev)
    q, k, v = _qkv(B, T, H, KV, dev, 1)
    plan = ops.AttnPlan(doc)
    valid = (doc > 0).reshape(-1)

    # --- row-sum property ---
    ones = torch.ones_like(v)
    o1, _ = ops.attn_fwd(q, k, ones, plan, H, KV, SCALE)
    assert torch.all(o1[~valid] == 0)
    assert float((o1[valid].float() - 1).abs().max()) <= 2 ** -7           # one bf16 ulp at 1.0

    # --- determinism + document independence (forward and backward) ---
    o, lse = ops.attn_fwd(q, k, v, plan, H, KV, SCALE)
    o_again, lse_again = ops.attn_fwd(q, k, v, plan, H, KV, SCALE)
    assert torch.equal(o, o_again) and torch.equal(lse, lse_again)
    do = torch.randn(B * T, H * 128, device=dev).bfloat16()
    dq, dk, dv = ops.attn_bwd(q, k, v, o, do, lse, plan, H, KV, SCALE)
    dq2, dk2, dv2 = ops.attn_bwd(q, k, v, o, do, lse, plan, H, KV, SCALE)
    assert torch.equal(dq, dq2) and torch.equal(dk, dk2) and torch.equal(dv, dv2)
    # linearity in dO: a power-of-two factor is exact in every intermediate (no denormals at these magnitudes)
    dq4, dk4, dv4 = ops.attn_bwd(q, k, v, o, (do.float() * 4).bfloat16(), lse, plan, H, KV, SCALE)
    assert torch.equal(dq4.float(), dq.float() * 4) and torch.equal(dv4.float(), dv.float() * 4)
    assert torch.equal(dk4.float(), dk.float() * 4)

    b, which = 0, len(lens[0]) // 2                                         # a document in the middle of row 0
    start = sum(lens[b][:which]); n = lens[b][which]
    Tn = (n + 127) // 128 * 128
    rows = slice(b * T + start, b * T + start + n)
    pad = lambda x: torch.cat([x[rows], torch.zeros(Tn - n, x.shape[1], device=dev, dtype=x.dtype)])
    doc1, _ = packed_doc_ids(1, Tn, [[n]], dev)
    plan1 = ops.AttnPlan(doc1)
    qa, ka, va, doa = pad(q), pad(k), pad(v), pad(do)
    oa, lsea = ops.attn_fwd(qa, ka, va, plan1, H, KV, SCALE)
    assert rel_err(oa[:n].float(), o[rows].float()) < 1e-2
    lse_row = lse.view(B, H, T)[b, :, start:start + n]
    assert float((lsea.view(1, H, Tn)[0, :, :n] - lse_row).abs().max()) < 2e-3
    dqa, dka, dva = ops.attn_bwd(qa, ka, va, oa, doa, lsea, plan1, H, KV, SCALE)
    for mine, alone in ((dq[rows], dqa[:n]), (dk[rows], dka[:n]), (dv[rows], dva[:n])):
        assert rel_err(alone.float(), mine.float()) < 2e-2

    # --- causality: perturb K/V of the last 100 positions of that document ---
    k2, v2 = k.clone(), v.clone()
    tail = slice(b * T + start + n - 100, b * T + start + n)
    k2[tail] = (k2[tail].float() * -1.5 + 0.25).bfloat16()
    v2[tail] = (v2[tail].float() + 3).bfloat16()
    o2, lse2 = ops.attn_fwd(q, k2, v2, plan, H, KV, SCALE)
    keep = torch.ones(B * T, dtype=torch.bool, device=dev)
    keep[tail] = False
    assert torch.equal(o2[keep], o[keep]), "rows outside the perturbed tail must not change by a single bit"
    assert not torch.equal(o2[tail], o[tail])


def test_single_document_32k_causal():
    """cfg 4 sequence length in one piece: T=32768, one document (maximum size; 17.6 TFLOP of attention forward)."""
    dev = require_cuda()
    B, T, H, KV = 1, 32768, 32, 8
    doc = torch.ones(B, T, dtype=torch.int64, device=dev)
    q, k, v = _qkv(B, T, H, KV, dev, 3)
    plan = ops.AttnPlan(doc)
    o1, lse = ops.attn_fwd(q, k, torch.ones_like(v), plan, H, KV, SCALE)
    assert float((o1.float() - 1).abs().max()) <= 2 ** -7
    assert torch.isfinite(lse).all()
    # first row attends only to itself: O[0] == V[0] exactly, lse[0] == scale * q0.k0
    o, lse = ops.attn_fwd(q, k, v, plan, H, KV, SCALE)
    v0 = v[0].view(KV, 128).repeat_interleave(H // KV, 0).reshape(-1)
    assert torch.equal(o[0], v0)
    s00 = (q[0].float().view(H, 128) * k[0].float().view(KV, 128).repeat_interleave(H // KV, 0)).sum(-1) * SCALE
    assert float((lse.view(H, T)[:, 0] - s00).abs().max()) < 2e-3
    # causality at the far end: changing the last K/V row changes only the last output row
    k2, v2 = k.clone(), v.clone()
    k2[-1] = -k2[-1]; v2[-1] = (v2[-1].float() + 5).bfloat16()
    o2, _ = ops.attn_fwd(q, k2, v2, plan, H, KV, SCALE)
    assert torch.equal(o2[:-1], o[:-1]) and not torch.equal(o2[-1], o[-1])
    do = torch.randn(B * T, H * 128, device=dev).bfloat16()
    dq, dk, dv = ops.attn_bwd(q, k, v, o, do, lse, plan, H, KV, SCALE)
    assert torch.isfinite(dq.float()).all() and torch.isfinite(dk.float()).all() and torch.isfinite(dv.float()).all()
    # dV of the LAST position receives only the last query's gradient: dV[T-1] = P[T-1,T-1] * dO[T-1] summed over the group
    p_last = torch.exp((q[-1].float().view(H, 128) * k[-1].float().view(KV, 128).repeat_interleave(H // KV, 0)).sum(-1) * SCALE
                       - lse.view(H, T)[:, -1])
    dv_last = (p_last[:, None] * do[-1].float().view(H, 128)).view(KV, H // KV, 128).sum(1).reshape(-1)
    assert rel_err(dv[-1].float(), dv_last) < 2e-2
