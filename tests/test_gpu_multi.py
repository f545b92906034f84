"""Multi-GPU parity: needs >= 2 B200s on the box, skipped otherwise (the driver's 1-GPU `-m gpu` run).
The N>1 host logic is also covered on CPU by tests/test_dist_gloo.py (gloo, world_size 2)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _torchrun(script, nproc, port, extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, script)]
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_context_parallel_matches_unsharded():
    """cp=2 (BASELINE config 4 shape of sharding): logits and all parameter gradients equal the unsharded run."""
    r = _torchrun("tools/check_cp.py", 2, 29541)
    assert r.returncode == 0 and "CP OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("fp32_params", ["0", "1"])
def test_fsdp2_matches_single_gpu(fp32_params):
    """dp_shard=2 (FSDP2, ref: touchnet/models/helper_func.py:134-202): same loss and gradients as one GPU on the
    concatenated batch, and the same loss after one SGD step (bf16 and fp32 all-gather dtypes)."""
    r = _torchrun("tools/check_fsdp.py", 2, 29542 + int(fp32_params), {"FSDP_FP32": fp32_params})
    assert r.returncode == 0 and "FSDP OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("audio", ["0", "1"])
def test_tensor_parallel_matches_unsharded(audio):
    """tp=2 with sequence parallelism (the reference's plan, parallelize_llama.py:105-196): logits and every parameter
    gradient equal the unsharded run; audio=1 adds the projector, q/k/v bias and the vocabulary-parallel embedding add."""
    r = _torchrun("tools/check_tp.py", 2, 29545 + int(audio), {"TP_AUDIO": audio})
    assert r.returncode == 0 and "TP OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
