"""GPU check of the peer-memory pull kernels (csrc/collective.cu) on ONE device: the "peers" are separate local buffers,
which exercises exactly the arithmetic and indexing the kernels perform on peer-mapped pointers.  Bit-exact: the
reduce-scatter sums in rank order in fp32, the all-gather is a copy.  (The FSDP2 plumbing around them is tested on CPU in
tests/test_parallel_gloo.py; the NVLink path itself has not run yet - see DESIGN.md 5.)"""
import pytest
import torch

from tests.gpu_util import require_cuda
from touchnet_b200 import fsdp_comm

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_peers,shard", [(2, 1 << 20), (4, 4096 + 8), (8, 1000), (3, 516)])
@pytest.mark.parametrize("avg", [False, True])
def test_peer_reduce_scatter(n_peers, shard, avg):
    dev = require_cuda()
    g = torch.Generator(device="cpu").manual_seed(n_peers * 7 + shard)
    ins = [torch.randn(n_peers * shard, generator=g).to(dev) for _ in range(n_peers)]
    scale = 1.0 / n_peers if avg else 1.0
    for rank in range(n_peers):
        for numel in (shard, shard - 3):                  # shard - 3: exercises the scalar tail (numel % 4 != 0)
            out = torch.full((shard,), float("nan"), device=dev)
            fsdp_comm._launch_reduce_scatter([t.data_ptr() for t in ins], rank * shard, out, numel, scale, 32)
            acc = ins[0][rank * shard:rank * shard + numel].clone()
            for t in ins[1:]:
                acc += t[rank * shard:rank * shard + numel]
            assert torch.equal(out[:numel], acc * scale), (n_peers, shard, rank, numel)
            assert torch.isnan(out[numel:]).all()          # nothing written past the requested range


@pytest.mark.parametrize("n_peers,n", [(2, 1 << 20), (4, 4096 + 8), (8, 1000 * 8)])
def test_peer_all_gather_in_place_layout(n_peers, n):
    """FSDP2's layout: every rank's shard sits at offset rank*n of its OWN full-size buffer."""
    dev = require_cuda()
    bufs = [torch.zeros(n_peers * n, dtype=torch.bfloat16, device=dev) for _ in range(n_peers)]
    shards = [torch.randn(n, device=dev).bfloat16() for _ in range(n_peers)]
    for p in range(n_peers):
        bufs[p][p * n:(p + 1) * n] = shards[p]
    es = 2
    for rank in range(n_peers):
        out = bufs[rank]
        fsdp_comm._launch_all_gather([bufs[p].data_ptr() + p * n * es for p in range(n_peers)], n * es, out, 32)
        assert torch.equal(out, torch.cat(shards)), (n_peers, n, rank)
