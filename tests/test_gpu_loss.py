"""GPU parity: fused pack-loss cross-entropy / accuracy vs the reference formulas (torch fp32 on the same bf16 logits).
ref: touchnet/loss/cross_entropy.py:12-50, touchnet/utils/metrics.py:26-50.  Tolerances: loss rel 1e-4;
dlogits max err <= 2^-8 relative to the row scale (bf16 output); accuracy / argmax exact."""
import pytest
import torch
import torch.nn.functional as F

from tests.gpu_util import require_cuda
from touchnet_b200 import loss as tl

pytestmark = pytest.mark.gpu


def ref_loss(pred, labels, sentence_lens, num_sentence):
    B = pred.shape[0]
    ce = F.cross_entropy(pred.flatten(0, 1).float(), labels.flatten(0, 1), reduction="none", ignore_index=-100)
    num_tokens = (labels != -100).sum()
    lpt = ce.sum() / num_tokens if (ce.sum() > 1e-6 and num_tokens > 0) else torch.zeros_like(ce.sum())
    lps = torch.sum(torch.sum(ce.reshape(B, -1) / sentence_lens, dim=-1)) / num_sentence
    return lps, lpt


@pytest.mark.parametrize("B,T,V", [(2, 256, 1000), (1, 512, 128256), (3, 77, 156032), (2, 64, 1003)])
def test_pack_ce_matches_reference(B, T, V):
    dev = require_cuda()
    torch.manual_seed(B * T + V)
    if V % 8 != 0:
        base = torch.randn(B, T, V + (8 - V % 8), device=dev).bfloat16()
        logits = base[..., :V]                                       # row stride stays 16-byte aligned, V has a tail
    else:
        logits = (torch.randn(B, T, V, device=dev) * 2).bfloat16()
    labels = torch.randint(0, V, (B, T), device=dev)
    labels[:, -T // 4:] = -100
    labels[0, :3] = -100
    sl = torch.randint(1, 50, (B, T), device=dev)
    ns = 7
    lg_ref = logits.detach().float().requires_grad_(True)
    lps_ref, lpt_ref = ref_loss(lg_ref, labels, sl, ns)
    lps_ref.backward()
    acc_ref = ((lg_ref.argmax(-1) == labels) & (labels != -100)).sum() / (labels != -100).sum()

    lg = logits.detach().clone().requires_grad_(True) if V % 8 == 0 else logits.detach().requires_grad_(True)
    lps, lpt = tl.cross_entropy_loss(lg, labels, sl, ns)
    acc = tl.accuracy(lg, labels)
    assert abs(float(lps) - float(lps_ref)) <= 1e-4 * abs(float(lps_ref)) + 1e-6
    assert abs(float(lpt) - float(lpt_ref)) <= 1e-4 * abs(float(lpt_ref)) + 1e-6
    assert float(acc) == pytest.approx(float(acc_ref), abs=1e-7)
    (lps * 3.0).backward()
    g, g_ref = lg.grad.float(), lg_ref.grad * 3.0
    scale = g_ref.abs().amax(dim=-1, keepdim=True).clamp(min=1e-12)
    assert float(((g - g_ref).abs() / scale).max()) <= 2 ** -7
    assert torch.all(g[labels == -100] == 0)


def test_argmax_ties_pick_first_like_torch():
    dev = require_cuda()
    logits = torch.zeros(1, 4, 1024, device=dev, dtype=torch.bfloat16)
    logits[0, 0, 5] = 1; logits[0, 0, 900] = 1        # tie: torch.argmax returns the first
    logits[0, 1, 1023] = 3
    labels = torch.tensor([[5, 1023, 0, -100]], device=dev)
    acc = tl.accuracy(logits, labels)
    assert float(acc) == pytest.approx(float(((logits.float().argmax(-1) == labels) & (labels != -100)).sum() / 3))
