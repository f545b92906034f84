"""CPU: structural drop-in contract of the modules (no compute): state-dict keys identical to HF LlamaForCausalLM /
the reference's TouchAudioForCausalLM, meta-device construction, post_init attributes the reference touches
(ref: touchnet/models/llama/__init__.py:19-36, touchnet/bin/train.py:179-182,274-283), FSDP-wrappable block list."""
import json
import os

import numpy as np
import pytest
import torch

from touchnet_b200 import modeling


class _Cfg:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _cfg(tie=False):
    return _Cfg(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=1,
                head_dim=128, vocab_size=64, rms_norm_eps=1e-5, rope_theta=500000.0,
                rope_scaling={"rope_type": "llama3", "factor": 32.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                              "original_max_position_embeddings": 8192},
                attention_bias=False, tie_word_embeddings=tie, initializer_range=0.02, model_type="llama")


def test_state_dict_keys_match_hf_llama():
    from transformers import LlamaConfig, LlamaForCausalLM
    c = _cfg()
    with torch.device("meta"):
        hf = LlamaForCausalLM(LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                                          num_attention_heads=2, num_key_value_heads=1, head_dim=128, vocab_size=64,
                                          tie_word_embeddings=False))
        ours = modeling.B200LlamaForCausalLM(c)                   # constructible on the meta device (train.py:179-182)
    hk = {k: tuple(v.shape) for k, v in hf.state_dict().items()}
    ok = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    assert hk == ok


def test_touch_audio_keys_match_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "model_tiny.npz"))
    ref_keys = sorted(k[6:] for k in g.files if k.startswith("ta_sd/"))
    cj = json.loads(bytes(g["config_json"]).decode())
    text = _Cfg(**{**cj, "head_dim": 128, "hidden_size": 1024})   # head_dim 128 variant of the reference's tiny config
    text.tie_word_embeddings = cj["tie_word_embeddings"]
    with torch.device("meta"):
        m = modeling.B200TouchAudioForCausalLM(_Cfg(audio_config=_Cfg(input_size=40), text_config=text, pad_token_id=0))
    ours = sorted(m.state_dict().keys())
    # tied lm_head: HF omits nothing from state_dict either; compare as sets of names
    assert set(ours) == set(ref_keys)


def test_post_init_contract_and_reference_post_init_runs():
    c = _cfg()
    m = modeling.B200LlamaForCausalLM(c)
    m.post_init()
    rot = m.model.rotary_emb
    inv, scaling = rot.rope_init_fn(rot.config, device=torch.device("cpu"))      # exactly what the reference calls
    assert torch.equal(inv, rot.inv_freq) and scaling == 1.0
    for attr in ("config", "rope_init_fn", "inv_freq", "attention_scaling", "original_inv_freq"):
        assert hasattr(rot, attr)
    for layer in m.model.layers:
        assert torch.all(layer.input_layernorm.weight == 1) and torch.all(layer.post_attention_layernorm.weight == 1)
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            assert isinstance(getattr(layer.self_attn, n), torch.nn.Linear)
        for n in ("gate_proj", "up_proj", "down_proj"):
            assert isinstance(getattr(layer.mlp, n), torch.nn.Linear)
    assert m.base_model_prefix == "model" and isinstance(m.model.layers, torch.nn.ModuleList)
    # restatement of the reference's get_num_params(exclude_embedding=True) contract (models/llama/__init__.py:57-67)
    n_all = sum(p.numel() for p in m.parameters())
    n_emb = sum(sum(p.numel() for p in mod.parameters()) for mod in m.model.children() if isinstance(mod, torch.nn.Embedding))
    assert n_all - n_emb == n_all - 64 * 256


def test_tied_embeddings_share_storage():
    m = modeling.B200LlamaForCausalLM(_cfg(tie=True))
    assert m.lm_head.weight is m.model.embed_tokens.weight and m._tied_weights_keys == ["lm_head.weight"]


def test_rope_parameters_match_golden_inv_freq(golden_dir):
    g = np.load(os.path.join(golden_dir, "model_tiny.npz"))
    cj = json.loads(bytes(g["config_json"]).decode())
    inv, _ = modeling.compute_rope_parameters(_Cfg(**cj), device=None)
    assert torch.equal(inv, torch.from_numpy(g["llama/inv_freq"]))   # HF llama3 rope init, bit exact


def test_head_dim_other_than_128_is_padded_or_refused_loudly():
    """head_dim 64 (the reference's example config Llama-3_2-1B.json:13) is accepted - it runs through the 128-wide attention
    kernels with zero-padded heads (tests/test_gpu_model.py) - anything the kernels cannot express is refused at construction."""
    c = _cfg()
    c.head_dim = 64
    m = modeling.B200LlamaForCausalLM(c)
    a = m.model.layers[0].self_attn
    assert a.head_dim == 64 and a.q_proj.weight.shape[0] == a.num_heads * 64
    for bad in (256, 60):
        c.head_dim = bad
        with pytest.raises(Exception, match="head_dim 128"):
            modeling.B200LlamaForCausalLM(c)


def test_cpu_tensors_are_refused_loudly():
    from touchnet_b200 import ops
    with pytest.raises(Exception, match="no CPU path|CUDA"):
        ops.rmsnorm_fwd(torch.zeros(4, 64, dtype=torch.bfloat16), torch.ones(64), 1e-5)


def test_public_modules_import_without_cuda_and_specs_carry_the_b200_parallelize_fn():
    """Every module of the package imports on a CPU-only host (the native library loads lazily), and the "*_b200" specs
    route parallelisation through touchnet_b200.parallelize (TP / CP on the fused block) instead of the reference's
    DTensor module plan."""
    import importlib
    import touchnet_b200
    for name in touchnet_b200.__all__:
        importlib.import_module(f"touchnet_b200.{name}")
    from touchnet_b200 import train_spec
    names = train_spec.register()
    assert set(names) == {"llama_b200", "touch_audio_b200"}
    try:
        spec = train_spec.get_train_spec("llama_b200")
    except ValueError:
        from touchnet.utils.train_spec import get_train_spec          # reference importable: its registry holds the spec
        spec = get_train_spec("llama_b200")
    assert spec.parallelize_fn.__name__ == "parallelize_b200"
    # the loss / accuracy slots the train loop consumes (ref: touchnet/bin/train.py:447-450) are the CUDA pack-loss, not
    # the reference's compiled fp32-upcast cross-entropy
    from touchnet_b200 import loss
    assert spec.loss_fn is loss.cross_entropy_loss and spec.acc_fn is loss.accuracy


def test_parallelize_fn_hands_the_reference_function_a_view_without_tp():
    """With the reference's own parallelize function available it still applies AC / FSDP2, but must not re-apply its
    DTensor tensor-parallel plan to the fused block: it sees parallel_dims with tp_enabled == False."""
    from types import SimpleNamespace
    import torch
    from touchnet_b200 import parallelize
    seen = {}

    def base_fn(model, world_mesh, dims, job_config):
        seen.update(tp=dims.tp_enabled, dp=dims.dp_shard_enabled, loss_parallel=dims.loss_parallel_enabled)
        return model

    dims = SimpleNamespace(tp_enabled=False, cp_enabled=False, pp_enabled=False, dp_shard_enabled=True,
                           dp_replicate_enabled=False, loss_parallel_enabled=False)
    m = torch.nn.Linear(2, 2)
    out = parallelize.make_parallelize_fn(base_fn)(m, None, dims, SimpleNamespace())
    assert out is m and seen == {"tp": False, "dp": True, "loss_parallel": False}
    view = parallelize._WithoutTP(SimpleNamespace(tp_enabled=True, tp=2))
    assert view.tp_enabled is False and view.tp == 2
    import pytest
    with pytest.raises(NotImplementedError):
        parallelize.make_parallelize_fn(None)(m, None, SimpleNamespace(pp_enabled=True), SimpleNamespace())
