"""CPU, world_size 2 over gloo: the N>1 bookkeeping of the sharded path (per-rank rows, max-over-ranks timing, whole-job
token accounting) and that two ranks really draw different packed rows with the reference layout."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from touchnet_b200 import batching, dist_util


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        seed = dist_util.rank_seed(2025, rank)
        buf, placed = batching.plan_audio_text_batch(seed, 1, 1024, 1000, stride=4, max_s=5.0)
        doc = buf["attention_mask"]
        sig = int(doc.sum()) * 1000 + len(placed)
        fake_ms = 10.0 + 5.0 * rank                                  # rank 1 is the slow one
        ms = dist_util.max_over_ranks(fake_ms, "cpu")
        total_nonpad = dist_util.sum_over_ranks(float((doc > 0).sum()), "cpu")
        tps = dist_util.whole_job_tokens_per_s(1024, 3, world, ms)
        gathered = [None] * world
        dist.all_gather_object(gathered, sig)
        q.put((rank, ms, tps, total_nonpad, gathered, int((doc > 0).sum())))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, ms0, tps0, tot0, g0, n0), (r1, ms1, tps1, tot1, g1, n1) = res
    assert ms0 == ms1 == 15.0                                        # max over ranks, identical on every rank
    assert tps0 == tps1 == pytest.approx(1024 * 2 * 3 / 0.015)       # whole-job aggregate, pads included
    assert tot0 == tot1 == n0 + n1
    assert g0 == g1 and g0[0] != g0[1]                               # the two ranks hold different packed rows
