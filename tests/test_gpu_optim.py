"""GPU parity of the optimizer step next to the path through the C ABI:
tn_sumsq_f32 / tn_scale_f32 / tn_adamw_f32 behind touchnet_b200.optim vs torch.nn.utils.clip_grad_norm_ and
torch.optim.AdamW (for-loop implementation, fp32) on identical tensors.  Tolerance: fp32 rounding-order only -
|p - p_ref| <= 2e-6 * max|p_ref| per tensor after 3 steps, moments likewise; bf16 working copies bit-equal to
p.bfloat16()."""
import pytest
import torch

from tests.gpu_util import require_cuda
from touchnet_b200 import _lib, modeling, ops, optim

pytestmark = pytest.mark.gpu


def _close(a, b, tol=2e-6):
    scale = float(b.abs().max()) + 1e-30
    return float((a - b).abs().max()) <= tol * scale


def _make(dev, shapes, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return [torch.randn(s, generator=g).to(dev) for s in shapes]


SHAPES = [(257, 129), (1024, 512), (4099,), (3,), (64, 64)]


def test_sumsq_and_clip_match_torch():
    dev = require_cuda()
    ours = [torch.nn.Parameter(t.clone()) for t in _make(dev, SHAPES, 1)]
    ref = [torch.nn.Parameter(t.detach().clone()) for t in ours]
    for a, b, g in zip(ours, ref, _make(dev, SHAPES, 2)):
        a.grad, b.grad = (g * 3).clone(), (g * 3).clone()
    n_ref = torch.nn.utils.clip_grad_norm_(ref, max_norm=1.0, foreach=False)
    n_ours = optim.clip_grad_norm_(ours, max_norm=1.0)
    assert abs(float(n_ours) - float(n_ref)) <= 1e-6 * float(n_ref)
    for a, b in zip(ours, ref):
        assert _close(a.grad, b.grad)
    # below the threshold nothing changes (coefficient clamps to 1)
    before = [p.grad.clone() for p in ours]
    optim.clip_grad_norm_(ours, max_norm=1e9)
    for p, g0 in zip(ours, before):
        assert torch.equal(p.grad, g0)


@pytest.mark.parametrize("deferred_clip", [False, True])
def test_adamw_matches_torch(deferred_clip):
    dev = require_cuda()
    ours = [torch.nn.Parameter(t.clone()) for t in _make(dev, SHAPES, 3)]
    ref = [torch.nn.Parameter(t.detach().clone()) for t in ours]
    kw = dict(lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    o_ours = optim.B200AdamW(ours, fused=True, foreach=False, **kw)      # the reference passes fused / foreach along
    o_ref = torch.optim.AdamW(ref, foreach=False, fused=False, **kw)
    for step in range(3):
        for a, b, g in zip(ours, ref, _make(dev, SHAPES, 10 + step)):
            a.grad, b.grad = (g * 2).clone(), (g * 2).clone()
        torch.nn.utils.clip_grad_norm_(ref, max_norm=1.0, foreach=False)
        v0 = ours[0]._version
        if deferred_clip:
            g_before = ours[0].grad.clone()
            optim.clip_grad_norm_(ours, max_norm=1.0, defer_to=o_ours)
            assert torch.equal(ours[0].grad, g_before)                   # untouched: the step applies the coefficient
        else:
            optim.clip_grad_norm_(ours, max_norm=1.0)
        o_ours.step()
        o_ref.step()
        assert ours[0]._version > v0                                     # raw-pointer write is visible to autograd / caches
        for a, b in zip(ours, ref):
            assert _close(a.detach(), b.detach()), step
            assert _close(o_ours.state[a]["exp_avg"], o_ref.state[b]["exp_avg"])
            assert _close(o_ours.state[a]["exp_avg_sq"], o_ref.state[b]["exp_avg_sq"])
    assert set(o_ours.state[ours[0]]) == {"step", "exp_avg", "exp_avg_sq"}          # DCP-compatible state keys
    assert float(o_ours.state[ours[0]]["step"]) == 3.0


def test_step_refreshes_bf16_working_copies():
    """After optimizer.step() the next forward issues no fp32->bf16 weight cast: the step wrote the working copies."""
    dev = require_cuda()
    from tests.test_gpu_model import small_cfg
    from tests.gpu_util import packed_doc_ids
    cfg = small_cfg(L=2, d=256, H=2, KV=1)
    torch.manual_seed(0)
    model = modeling.B200LlamaForCausalLM(cfg).to(dev)
    model.post_init()
    opt = optim.B200AdamW(model.parameters(), lr=1e-2)
    B, T = 1, 256
    doc, pos = packed_doc_ids(B, T, [[200, 56]], dev)
    ids = torch.randint(0, cfg.vocab_size, (B, T), device=dev)

    def fwd_bwd():
        out = model(input_ids=ids, attention_mask=doc, position_ids=pos).logits
        out.float().square().mean().backward()
        return out.detach()

    names = []
    hook = lambda name, phase, args: names.append(name) if phase == "pre" else None
    y0 = fwd_bwd()
    optim.clip_grad_norm_(model.parameters(), 1.0, defer_to=opt)
    opt.step()
    opt.zero_grad()
    _lib._hooks.append(hook)
    try:
        y1 = fwd_bwd()
    finally:
        _lib._hooks.remove(hook)
    torch.cuda.synchronize()
    assert "tn_cast_f32_bf16" not in names, "the optimizer step should have refreshed every bf16 working copy"
    assert not torch.equal(y0, y1)                                        # the update took effect
    for m in model.modules():
        if isinstance(m, torch.nn.Linear):
            assert torch.equal(ops.bf16_weight(m.weight), m.weight.detach().bfloat16())
