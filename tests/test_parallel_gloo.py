"""CPU, world_size 2 over gloo: the host logic of tensor parallelism (touchnet_b200/tensor_parallel.py) and context
parallelism (touchnet_b200/context_parallel.py) - DTensor parameter sharding, the collectives inside the decoder block's
forward/backward, vocabulary-parallel embedding / lm_head, gradient reductions - against the unsharded run of the same
model.  The CUDA ops are replaced by plain-torch stand-ins (tests/cpu_ops_shim.py, test infrastructure only), so what is
compared is exactly the wiring; kernel numerics are covered by the `-m gpu` tests and tests/test_gpu_multi.py."""
import copy
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _Cfg:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _text_cfg(bias=False):
    return _Cfg(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                head_dim=128, vocab_size=384, rms_norm_eps=1e-5, rope_theta=500000.0, rope_scaling=None,
                attention_bias=bias, tie_word_embeddings=False, initializer_range=0.02, model_type="llama", pad_token_id=0)


def _docs(B, T, lens):
    doc = torch.zeros(B, T, dtype=torch.int64)
    pos = torch.zeros(B, T, dtype=torch.int64)
    for b, ls in enumerate(lens):
        o = 0
        for i, n in enumerate(ls):
            doc[b, o:o + n] = i + 1
            pos[b, o:o + n] = torch.arange(n)
            o += n
    return doc, pos


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _build(audio: bool, bias: bool):
    from touchnet_b200 import modeling
    torch.manual_seed(11)
    text = _text_cfg(bias)
    if audio:
        model = modeling.B200TouchAudioForCausalLM(_Cfg(audio_config=_Cfg(input_size=80), text_config=text, pad_token_id=0))
    else:
        model = modeling.B200LlamaForCausalLM(text)
    model.post_init()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 2:
                p.normal_(0, 0.05)
            elif "bias" in n:
                p.normal_(0, 0.1)
            else:
                p.uniform_(0.5, 1.5)                      # norm weights away from 1 so their gradients matter
    return model, text


def _inputs(B, T, V, audio):
    g = torch.Generator().manual_seed(5)
    doc, pos = _docs(B, T, [[100, 120, 20], [256]][:B])
    ids = torch.randint(1, V, (B, T), generator=g)
    tgt = torch.randn(B, T, V, generator=g)
    kw = dict(input_ids=ids, attention_mask=doc, position_ids=pos)
    if audio:
        is_audio = torch.zeros(B, T, dtype=torch.bool)
        is_audio[:, :64] = True
        kw["input_features"] = torch.randn(B, T, 80, generator=g) * is_audio[..., None]
        kw["input_ids"] = torch.where(is_audio, torch.zeros_like(ids), ids)
    return kw, doc, tgt


def _loss(logits, tgt, doc, denom):
    return ((logits.float() * tgt)[doc > 0]).sum() / denom


class FilePeerMemory:
    """Test double of tensor_parallel.SymmPeerMemory: /dev/shm files mapped by every rank stand in for symmetric memory
    (a store into views[peer] is visible to the peer after the barrier, like an NVLink P2P store)."""
    _count = 0

    def __init__(self, group, device):
        self.group = group
        self.size, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.paths = []

    def alloc(self, shape, dtype):
        FilePeerMemory._count += 1
        numel = 1
        for d in shape:
            numel *= d
        base = f"/dev/shm/tn_peer_{os.environ['MASTER_PORT']}_{FilePeerMemory._count}"
        mine = f"{base}_{self.rank}"
        torch.from_file(mine, shared=True, size=numel, dtype=dtype).zero_()          # creates the file
        dist.barrier(self.group)
        self.paths.append(mine)
        return [torch.from_file(f"{base}_{r}", shared=True, size=numel, dtype=dtype).view(*shape) for r in range(self.size)]

    def barrier(self):
        dist.barrier(self.group)

    def cleanup(self):
        for p in self.paths:
            try:
                os.remove(p)
            except OSError:
                pass


def _tp_worker(rank, world, port, audio, bias, q, peer=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests import cpu_ops_shim
        cpu_ops_shim.install()
        from torch.distributed.device_mesh import init_device_mesh
        from touchnet_b200 import tensor_parallel
        if peer:        # NCCL-free context: same code path as TN_TP_PEER=1 on GPUs, symmetric memory replaced by /dev/shm
            os.environ["TN_TP_PEER"] = "1"
            tensor_parallel.SymmPeerMemory = FilePeerMemory
        model, text = _build(audio, bias)
        B, T = 2, 256
        kw, doc, tgt = _inputs(B, T, text.vocab_size, audio)
        denom = float((doc > 0).sum()) * text.vocab_size
        ref_model = copy.deepcopy(model)
        ref_logits = ref_model(**kw).logits
        _loss(ref_logits, tgt, doc, denom).backward()
        ref_grads = {n: p.grad.clone() for n, p in ref_model.named_parameters()}

        mesh = init_device_mesh("cpu", (world,), mesh_dim_names=("tp",))
        tensor_parallel.apply_tp(model, mesh)
        sd = dict(model.named_parameters())
        assert set(sd) == set(ref_grads)                                  # FQNs unchanged by the sharding
        lm = "language_model." if audio else ""
        assert sd[lm + "model.layers.0.self_attn.q_proj.weight"].to_local().shape == (4 * 128 // world, 256)
        assert sd[lm + "model.layers.0.mlp.down_proj.weight"].to_local().shape == (256, 512 // world)
        assert sd[lm + "model.embed_tokens.weight"].to_local().shape == (384 // world, 256)
        logits = model(**kw).logits
        assert logits.shape == ref_logits.shape
        _loss(logits, tgt, doc, denom).backward()
        err_fwd = float((logits.float() - ref_logits.float())[doc > 0].abs().max()) / float(ref_logits.float().abs().max())
        worst, worst_name = 0.0, ""
        for n, p in model.named_parameters():
            g = p.grad
            g = g.full_tensor() if tensor_parallel.is_dtensor(g) else g
            e = _rel(g.float(), ref_grads[n].float())
            if e > worst:
                worst, worst_name = e, n
        if peer:
            ctx = (model.language_model.model if audio else model.model)._tn_tp_peer_ctx
            assert isinstance(ctx, tensor_parallel.PeerTPContext) and len(ctx._kept) == 4      # h1, h2 of 2 blocks
            ctx.mem.cleanup()
        q.put((rank, err_fwd, worst, worst_name))
    finally:
        dist.destroy_process_group()


def _tp_worker_peer(rank, world, port, q):
    _tp_worker(rank, world, port, True, True, q, peer=True)


def _cp_worker(rank, world, port, q, halo=False, head_tail=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests import cpu_ops_shim
        cpu_ops_shim.install()
        from touchnet_b200 import context_parallel
        model, text = _build(False, False)
        B, T = 2, 256
        kw, doc, tgt = _inputs(B, T, text.vocab_size, False)
        if halo or head_tail:   # longer rows: the halo is a strict subset of the neighbour's shard / four chunks of 256
            if halo:
                os.environ["TN_CP_HALO"] = "1"
            B, T = 2, 1024
            g = torch.Generator().manual_seed(9)
            doc, pos = _docs(B, T, [[300, 300, 300], [200, 400, 300]])
            kw = dict(input_ids=torch.randint(1, text.vocab_size, (B, T), generator=g), attention_mask=doc, position_ids=pos)
            tgt = torch.randn(B, T, text.vocab_size, generator=g)
        denom = float((doc > 0).sum()) * text.vocab_size
        ref_logits = model(**kw).logits
        _loss(ref_logits, tgt, doc, denom).backward()
        ref_grads = {n: p.grad.clone() for n, p in model.named_parameters()}
        model.zero_grad()
        Tl = T // world
        sl = slice(rank * Tl, (rank + 1) * Tl)
        if head_tail:           # torch's load-balanced layout: chunk r followed by chunk 2*cp-1-r
            Tw = Tl // 2
            sl = torch.cat([torch.arange(rank * Tw, (rank + 1) * Tw),
                            torch.arange((2 * world - 1 - rank) * Tw, (2 * world - rank) * Tw)])
        context_parallel.enable_context_parallel(model, dist.group.WORLD, load_balance=head_tail)
        if halo:
            plan = context_parallel.make_cp_plan(doc[:, sl].contiguous(), dist.group.WORLD)
            assert context_parallel.halo_first_blocks(plan) == [0, 1]      # rank 1 needs rows 128.. of rank 0, not 0..
            assert context_parallel._halo_pairs(plan) == [(0, 1, 128, 512)]
        lg = model(**{k: v[:, sl].contiguous() for k, v in kw.items()}).logits
        _loss(lg, tgt[:, sl], doc[:, sl], denom).backward()
        err_fwd = float((lg.float() - ref_logits[:, sl].float()).abs().max()) / float(ref_logits.float().abs().max())
        worst, worst_name = 0.0, ""
        for n, p in model.named_parameters():
            g = p.grad.clone()
            dist.all_reduce(g)                       # every rank holds the partial sum over its tokens
            e = _rel(g.float(), ref_grads[n].float())
            if e > worst:
                worst, worst_name = e, n
        q.put((rank, err_fwd, worst, worst_name))
    finally:
        dist.destroy_process_group()


def _tp_fsdp_worker(rank, world, port, q):
    """TP=2 x FSDP2=2 (BASELINE config 5 at a quarter of the mesh): 2-D DTensor parameters, each dp rank its own row."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests import cpu_ops_shim
        cpu_ops_shim.install()
        from torch.distributed.device_mesh import init_device_mesh
        from torch.distributed.fsdp import fully_shard
        from touchnet_b200 import tensor_parallel
        model, text = _build(False, False)
        B, T = 2, 256
        kw, doc, tgt = _inputs(B, T, text.vocab_size, False)
        denom = float((doc > 0).sum()) * text.vocab_size
        ref_model = copy.deepcopy(model)
        ref_logits = ref_model(**kw).logits
        _loss(ref_logits, tgt, doc, denom).backward()
        ref_grads = {n: p.grad.clone() for n, p in ref_model.named_parameters()}

        # through the TrainSpec parallelize_fn (touchnet_b200/parallelize.py), as touchnet/bin/train.py:184-190 calls it
        from touchnet_b200 import parallelize
        mesh = init_device_mesh("cpu", (2, 2), mesh_dim_names=("dp_shard_cp", "tp"))
        dims = _Cfg(tp_enabled=True, cp_enabled=False, pp_enabled=False, dp_shard_enabled=True, dp_replicate_enabled=False)
        job = _Cfg(training_mixed_precision_param="float32", training_mixed_precision_reduce="float32",
                   training_fsdp_reshard_after_forward="default")
        model = parallelize.make_parallelize_fn(None)(model, mesh, dims, job)
        d = mesh["dp_shard_cp"].get_local_rank()
        row = slice(d, d + 1)
        logits = model(**{k: v[row] for k, v in kw.items()}).logits
        _loss(logits, tgt[row], doc[row], denom).backward()
        err_fwd = float((logits.float() - ref_logits[row].float())[doc[row] > 0].abs().max()) / float(ref_logits.float().abs().max())
        worst, worst_name = 0.0, ""
        for n, p in model.named_parameters():
            g = p.grad.full_tensor() * 2                                   # FSDP averages over dp; rank losses are partial sums
            e = _rel(g.float(), ref_grads[n].float())
            if e > worst:
                worst, worst_name = e, n
        q.put((rank, err_fwd, worst, worst_name))
    finally:
        dist.destroy_process_group()


def _cp_worker_halo(rank, world, port, q):
    _cp_worker(rank, world, port, q, halo=True)


def _cp_worker_head_tail(rank, world, port, q):
    _cp_worker(rank, world, port, q, head_tail=True)


def _cp_fsdp_worker(rank, world, port, q):
    """CP=2 x FSDP2 over the flattened dp_shard_cp mesh of 4 (BASELINE config 4 at half the mesh;
    ref: touchnet/utils/distributed.py:116-157 builds the same flattened mesh for apply_fsdp)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests import cpu_ops_shim
        cpu_ops_shim.install()
        from torch.distributed.device_mesh import init_device_mesh
        from torch.distributed.fsdp import fully_shard
        from touchnet_b200 import context_parallel
        model, text = _build(False, False)
        B, T = 2, 256
        kw, doc, tgt = _inputs(B, T, text.vocab_size, False)
        denom = float((doc > 0).sum()) * text.vocab_size
        ref_model = copy.deepcopy(model)
        ref_logits = ref_model(**kw).logits
        _loss(ref_logits, tgt, doc, denom).backward()
        ref_grads = {n: p.grad.clone() for n, p in ref_model.named_parameters()}

        mesh = init_device_mesh("cpu", (2, 2), mesh_dim_names=("dp_shard", "cp"))
        flat = mesh["dp_shard", "cp"]._flatten("dp_shard_cp")
        context_parallel.enable_context_parallel(model, mesh["cp"].get_group())
        for layer in model.model.layers:
            fully_shard(layer, mesh=flat)
        fully_shard(model, mesh=flat)
        d, c = mesh["dp_shard"].get_local_rank(), mesh["cp"].get_local_rank()
        Tl = T // 2
        row, sl = slice(d, d + 1), slice(c * Tl, (c + 1) * Tl)
        logits = model(**{k: v[row, sl].contiguous() for k, v in kw.items()}).logits
        _loss(logits, tgt[row, sl], doc[row, sl], denom).backward()
        err_fwd = float((logits.float() - ref_logits[row, sl].float()).abs().max()) / float(ref_logits.float().abs().max())
        worst, worst_name = 0.0, ""
        for n, p in model.named_parameters():
            g = p.grad.full_tensor() * 4                                   # FSDP averages over the 4 ranks' partial sums
            e = _rel(g.float(), ref_grads[n].float())
            if e > worst:
                worst, worst_name = e, n
        q.put((rank, err_fwd, worst, worst_name))
    finally:
        dist.destroy_process_group()


def _fsdp_peer_worker(rank, world, port, mode, q):
    """FSDP2=2 with touchnet_b200.fsdp_comm's peer-memory collectives plugged into every module group: the Comm
    protocol (allocate / call), barrier placement and buffer ring against the unsharded model.  Symmetric memory ->
    /dev/shm files, the two pull kernels -> torch on the mapped buffers."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ctypes
        from tests import cpu_ops_shim
        cpu_ops_shim.install()
        from torch.distributed.device_mesh import init_device_mesh
        from torch.distributed.fsdp import MixedPrecisionPolicy, fully_shard
        from touchnet_b200 import fsdp_comm

        def at(ptr, n, dtype):
            es = torch.empty(0, dtype=dtype).element_size()
            return torch.frombuffer((ctypes.c_char * (n * es)).from_address(ptr), dtype=dtype)

        calls = {"rs": 0, "ag": 0}

        def fake_rs(ptrs, shard_offset, out, numel, scale, max_ctas):
            calls["rs"] += 1
            acc = at(ptrs[0] + 4 * shard_offset, numel, torch.float32).clone()
            for p in ptrs[1:]:
                acc += at(p + 4 * shard_offset, numel, torch.float32)
            out.view(-1).copy_(acc * scale)

        def fake_ag(ptrs, bytes_each, out, max_ctas):
            calls["ag"] += 1
            o = out.view(-1).view(torch.uint8)
            for i, p in enumerate(ptrs):
                o[i * bytes_each:(i + 1) * bytes_each] = at(p, bytes_each, torch.uint8)

        def fake_rs16(ptrs, out, numel, scale, max_ctas):                  # tn_reduce_bf16_to_f32: bf16 chunks, fp32 sum
            calls["rs"] += 1
            calls["rs16"] = calls.get("rs16", 0) + 1
            acc = at(ptrs[0], numel, torch.bfloat16).float()
            for p in ptrs[1:]:
                acc += at(p, numel, torch.bfloat16).float()
            out.view(-1).copy_(acc * scale)

        fsdp_comm._launch_reduce_scatter, fsdp_comm._launch_all_gather = fake_rs, fake_ag
        fsdp_comm._launch_reduce_bf16 = fake_rs16
        model, text = _build(False, False)
        B, T = 2, 256
        kw, doc, tgt = _inputs(B, T, text.vocab_size, False)
        denom = float((doc > 0).sum()) * text.vocab_size
        ref_model = copy.deepcopy(model)
        ref_logits = ref_model(**kw).logits
        _loss(ref_logits, tgt, doc, denom).backward()
        ref_grads = {n: p.grad.clone() for n, p in ref_model.named_parameters()}

        mesh = init_device_mesh("cpu", (world,), mesh_dim_names=("dp_shard",))
        mp_policy = MixedPrecisionPolicy(param_dtype=torch.bfloat16, reduce_dtype=torch.float32)
        for layer in model.model.layers:
            fully_shard(layer, mesh=mesh, mp_policy=mp_policy)
        fully_shard(model, mesh=mesh, mp_policy=mp_policy)
        mem = FilePeerMemory(mesh.get_group(), "cpu")
        pool = fsdp_comm.install(model, mesh.get_group(), "cpu", mem=mem, mode=mode)
        row = slice(rank, rank + 1)
        worst, worst_name, err_fwd = 0.0, "", 0.0
        for step in range(2):                                             # second step reuses the ring buffers
            model.zero_grad()
            logits = model(**{k: v[row] for k, v in kw.items()}).logits
            _loss(logits, tgt[row], doc[row], denom).backward()
            err_fwd = max(err_fwd, float((logits.float() - ref_logits[row].float())[doc[row] > 0].abs().max())
                          / float(ref_logits.float().abs().max()))
            for n, p in model.named_parameters():
                e = _rel(p.grad.full_tensor().float() * world, ref_grads[n].float())
                if e > worst:
                    worst, worst_name = e, n
        assert calls["rs"] == 2 * 3, calls                                # 2 blocks + root, every step
        assert calls["ag"] >= 2 * 3 if mode == "pull" else calls["ag"] == 0, calls   # push mode: copies only, no gather kernel
        if mode == "push":                                               # direct form: bf16 chunks straight from autograd's
            assert calls.get("rs16", 0) == 2 * 3, calls                   # gradients, FSDP2's copy-in bypassed every time
        assert all(len(r) <= fsdp_comm.RING for r in pool._rings.values())
        mem.cleanup()
        q.put((rank, err_fwd, worst, worst_name))
    finally:
        dist.destroy_process_group()


def _run(target, args, port_base, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = port_base + (os.getpid() % 150)
    procs = [ctx.Process(target=target, args=(r, world, port) + args + (q,)) for r in range(world)]
    for p in procs:
        p.start()
    res = []
    try:
        for _ in range(world):
            res.append(q.get(timeout=180))
    finally:
        for p in procs:
            p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    return sorted(res)


@pytest.mark.parametrize("audio,bias", [(False, False), (True, True)])
def test_tensor_parallel_matches_unsharded(audio, bias):
    for rank, err_fwd, worst, name in _run(_tp_worker, (audio, bias), 29700):
        assert err_fwd < 2e-2, (rank, err_fwd)       # bf16 rounding of partial sums before the reduce-scatter
        assert worst < 3e-2, (rank, name, worst)


def test_tensor_parallel_peer_memory_context_matches_unsharded():
    """tensor_parallel.PeerTPContext (opt-in, TN_TP_PEER=1): the block's collectives as direct stores into the peers'
    buffers - row-split GEMM launches whose outputs land in the destination rank's receive slot, gathers by peer stores,
    one barrier each.  Symmetric memory is replaced by shared-memory files; B=2 exercises the per-batch-row indexing."""
    for rank, err_fwd, worst, name in _run(_tp_worker_peer, (), 30340):
        assert err_fwd < 2e-2, (rank, err_fwd)
        assert worst < 3e-2, (rank, name, worst)


def test_context_parallel_matches_unsharded():
    for rank, err_fwd, worst, name in _run(_cp_worker, (), 29860):
        assert err_fwd < 1e-2, (rank, err_fwd)
        assert worst < 3e-2, (rank, name, worst)


def test_tensor_parallel_composes_with_fsdp2():
    for rank, err_fwd, worst, name in _run(_tp_fsdp_worker, (), 30020, world=4):
        assert err_fwd < 2e-2, (rank, err_fwd)
        assert worst < 3e-2, (rank, name, worst)


def test_context_parallel_halo_exchange_matches_unsharded():
    """TN_CP_HALO=1: only the K/V rows a rank's queries can reach are exchanged (here rows 128..511 of rank 0 instead of
    its whole shard), dK/dV of those rows travel back; results equal the unsharded run."""
    for rank, err_fwd, worst, name in _run(_cp_worker_halo, (), 30500):
        assert err_fwd < 1e-2, (rank, err_fwd)
        assert worst < 3e-2, (rank, name, worst)


def test_context_parallel_head_tail_layout_matches_unsharded():
    """load_balance=True: the sequence shards arrive in torch's head-tail order (rank r = chunk r + chunk 2cp-1-r, what
    `context_parallel` hands the model by default); two query windows per rank, K/V and dK/dV re-ordered around the
    collectives; results equal the unsharded run."""
    for rank, err_fwd, worst, name in _run(_cp_worker_head_tail, (), 30820):
        assert err_fwd < 1e-2, (rank, err_fwd)
        assert worst < 3e-2, (rank, name, worst)


def test_context_parallel_composes_with_fsdp2():
    for rank, err_fwd, worst, name in _run(_cp_fsdp_worker, (), 30180, world=4):
        assert err_fwd < 1e-2, (rank, err_fwd)
        assert worst < 3e-2, (rank, name, worst)


@pytest.mark.parametrize("mode", ["pull", "push"])
def test_fsdp2_peer_memory_collectives_match_unsharded(mode):
    """pull: tn_peer_* kernels read the peers' buffers; push: copy-engine pushes into the peers' buffers + local reduce."""
    for rank, err_fwd, worst, name in _run(_fsdp_peer_worker, (mode,), 30660 if mode == "pull" else 30700):
        assert err_fwd < 2e-2, (rank, err_fwd)
        assert worst < 3e-2, (rank, name, worst)
