"""Host-side contract of touchnet_b200.optim (no GPU): constructor mirrors what the reference builds
(ref: touchnet/utils/optimizer.py:153-166), and there is no CPU arithmetic path."""
import pytest
import torch

from touchnet_b200 import _lib, optim


def test_adamw_defaults_and_kwargs_follow_the_reference():
    p = torch.nn.Parameter(torch.zeros(4, 4))
    o = optim.B200AdamW([p], lr=8e-5, eps=1e-8, betas=(0.9, 0.95), weight_decay=0.1, fused=True, foreach=False)
    g = o.param_groups[0]
    assert (g["lr"], g["betas"], g["eps"], g["weight_decay"]) == (8e-5, (0.9, 0.95), 1e-8, 0.1)
    assert o.state_dict()["state"] == {}
    with pytest.raises(ValueError):
        optim.B200AdamW([p], lr=-1.0)


def test_no_cpu_path():
    p = torch.nn.Parameter(torch.zeros(8))
    p.grad = torch.ones(8)
    with pytest.raises(_lib.TouchNetB200Error):
        optim.clip_grad_norm_([p], 1.0)
    with pytest.raises(_lib.TouchNetB200Error):
        optim.clip_grad_norm_([p], 1.0, norm_type=float("inf"))
