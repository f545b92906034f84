"""CPU tests of the round-2 host logic (no kernel is launched): the lazy / vocabulary-parallel logits handles, the specs'
fused-loss model classes, the FSDP2-safety probe of the fused loss, the direct-push eligibility rule, the work-list
arithmetic of the GEMM tail split and the bench's tile accounting."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from touchnet_b200 import loss as tn_loss
from touchnet_b200 import modeling, train_spec


def test_logits_handles_carry_the_shape_contract_of_real_logits():
    """`pred.logits` of the "*_b200" specs in training mode: not a tensor, but it answers what the train loop / metrics code
    may ask of logits before handing them to loss_fn / acc_fn (ref: touchnet/bin/train.py:439-450)."""
    h = torch.zeros(2, 16, 8, dtype=torch.bfloat16)
    w = torch.zeros(100, 8)
    lz = tn_loss.LazyLogits(h, w)
    assert tuple(lz.shape) == (2, 16, 100) and lz.dim() == 3 and lz.size(-1) == 100 and lz.dtype == torch.bfloat16
    assert "materialize" in repr(lz)
    vp = tn_loss.VocabParallelLogits(torch.zeros(2, 16, 25, dtype=torch.bfloat16), group=None, v0=50, v_total=100)
    assert tuple(vp.shape) == (2, 16, 100) and vp.v0 == 50 and vp.dim() == 3
    # the fused path has no CPU form: it fails loudly instead of silently materialising
    with pytest.raises(Exception, match="CUDA|no CPU"):
        tn_loss.fused_linear_cross_entropy(h, w, torch.zeros(32, dtype=torch.int64), torch.ones(32, dtype=torch.int64), 1.0)


def test_spec_model_classes_switch_the_fused_loss_on_and_plain_classes_do_not():
    from types import SimpleNamespace as NS
    text = NS(hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=1, num_key_value_heads=1,
              head_dim=64, vocab_size=32, rms_norm_eps=1e-5, rope_theta=10000.0, rope_scaling=None, attention_bias=False,
              tie_word_embeddings=False, initializer_range=0.02, model_type="llama", pad_token_id=0)
    assert modeling.B200LlamaForCausalLM(text).fused_linear_ce is False
    assert train_spec.B200LlamaForCausalLMFused(text).fused_linear_ce is True
    cfg = NS(audio_config=NS(input_size=40), text_config=text, pad_token_id=0)
    assert modeling.B200TouchAudioForCausalLM(cfg).fused_linear_ce is False
    m = train_spec.B200TouchAudioForCausalLMFused(cfg)
    assert m.fused_linear_ce is True and m.language_model.fused_linear_ce is True
    m.fused_linear_ce = False                                   # the wrapper's property writes through
    assert m.language_model.fused_linear_ce is False
    assert m.language_model._tn_fsdp_root is m                  # lm_head belongs to the wrapper's (root) FSDP2 group
    assert modeling.B200LlamaForCausalLM(text).loss_parallel is False


def _reshard_probe_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from torch.distributed.device_mesh import init_device_mesh
        from torch.distributed.fsdp import fully_shard
        mesh = init_device_mesh("cpu", (world,), mesh_dim_names=("dp_shard",))
        res = []
        for reshard in (True, False):
            m = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.Linear(8, 8))
            fully_shard(m, mesh=mesh, reshard_after_forward=reshard)
            res.append(modeling._root_reshards_after_forward(m))
        res.append(modeling._root_reshards_after_forward(torch.nn.Linear(4, 4)))     # not FSDP-managed at all
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_fused_loss_is_only_used_when_the_fsdp2_root_keeps_its_parameters():
    """lm_head.weight is consumed by loss_fn AFTER forward returned: under FSDP2 that is only legal when the root group does
    not reshard after forward (the reference's default policy does, ref: touchnet/models/helper_func.py:196-202)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31100 + (os.getpid() % 200)
    procs = [ctx.Process(target=_reshard_probe_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    for rank, (resharding, keeping, plain) in res:
        assert resharding is True and keeping is False and plain is False, (rank, resharding, keeping, plain)


def test_direct_push_eligibility_rule():
    """PushReduceScatter pushes chunks straight from autograd's gradients only when that is the same arithmetic as FSDP2's
    chunk_cat copy-in: contiguous tensors of one dtype (bf16 / fp32) whose dim 0 splits evenly and whose chunks are 16-byte
    multiples; everything else takes torch's staged path."""
    from touchnet_b200 import fsdp_comm
    comm = fsdp_comm.PushReduceScatter.__new__(fsdp_comm.PushReduceScatter)
    comm.pool = type("P", (), {"size": 4, "rank": 0})()
    ok = [torch.zeros(16, 32, dtype=torch.bfloat16), torch.zeros(64, dtype=torch.bfloat16)]
    assert comm.can_direct(ok, 4)
    assert not comm.can_direct(ok, 2)                                              # another world size than the pool's
    assert not comm.can_direct([torch.zeros(18, 32, dtype=torch.bfloat16)], 4)      # dim 0 does not split evenly
    assert not comm.can_direct([torch.zeros(16, 32, dtype=torch.bfloat16).t()], 4)  # not contiguous
    assert not comm.can_direct(ok + [torch.zeros(16, dtype=torch.float32)], 4)      # mixed dtypes
    assert not comm.can_direct([torch.zeros(16, 3, dtype=torch.bfloat16)], 4)       # 4 x 3 x 2 B chunks: not 16-byte multiples
    assert not comm.can_direct([torch.zeros(16, 32, dtype=torch.float16)], 4)
    assert not comm.can_direct([], 4)
    comm.pool = type("P", (), {"size": 1, "rank": 0})()          # 1-rank group: FSDP2 never calls the reduce-scatter, so the
    assert not comm.can_direct(ok, 1)                            # copy-in must not be skipped


def _pair_work(w, num_tiles, C, allow=True):
    """Python restatement of csrc/gemm2.cu::pair_work / pair_num_work (documentation of the schedule, checked for the shapes
    of the Llama-3-8B step)."""
    full_w = (num_tiles // C) * C
    R = num_tiles - full_w
    split = allow and R > 0 and 2 * R <= C
    n = full_w + 2 * R if split else num_tiles
    if not split or w < full_w:
        return n, (w, -1)
    return n, (full_w + (w - full_w) // 2, (w - full_w) & 1)


@pytest.mark.parametrize("tiles,C,expect_split", [(256, 74, True), (384, 74, True), (896, 74, True), (1792, 74, True),
                                                  (512, 74, False), (768, 74, True), (8, 8, False), (16032, 74, False)])
def test_gemm_tail_split_covers_every_tile_exactly_once(tiles, C, expect_split):
    n, _ = _pair_work(0, tiles, C)
    assert (n > tiles) == expect_split
    seen = {}
    for w in range(n):
        _, (t, half) = _pair_work(w, tiles, C)
        seen.setdefault(t, []).append(half)
    assert sorted(seen) == list(range(tiles))
    for t, halves in seen.items():
        assert halves == [-1] or sorted(halves) == [0, 1], (t, halves)
    if expect_split:      # the split tail fits in ONE wave of half-length items
        assert n - (tiles // C) * C <= C


def test_bench_tile_accounting_matches_the_kernels_block_ranges():
    """bench.py's tile-granular attention FLOPs: (q block, kv block) pairs from the first block of the earliest document that
    reaches into a q block up to the diagonal - 219 pairs per head on the bench's seed-2025 row (what tn_attn_prep computes)."""
    import bench
    from touchnet_b200 import batching
    doc = batching.plan_audio_text_batch(2025, 1, 8192, 128256, stride=4, max_s=30.0)[0]["attention_mask"]
    assert bench.attn_tile_flops_fwd_per_layer(doc, H=1) == 219 * 4.0 * 128 * 128 * 128
    one = torch.ones(1, 512, dtype=torch.int64)
    assert bench.attn_tile_flops_fwd_per_layer(one, H=1) == (1 + 2 + 3 + 4) * 4.0 * 128 ** 3
    pad = torch.zeros(1, 256, dtype=torch.int64)
    assert bench.attn_tile_flops_fwd_per_layer(pad, H=1) == 0
