"""GPU parity: the drop-in modules (touchnet_b200/modeling.py) vs the oracle restatement of the HF forward on identical
weights and inputs, forward and backward, plus the structural contract (FQNs / state-dict keys / post_init attributes).

Tolerances (bf16 compute vs the oracle run in bf16 with the HF rounding points, and vs the fp32 oracle):
  logits: max |err| <= 3e-2 * max|ref| ; argmax identical wherever the oracle's top-2 margin exceeds 2x that error
  parameter gradients: relative L2 error < 3e-2 per tensor vs fp32 autograd of the oracle."""
import pytest
import torch

from oracle import model_oracle as mo
from tests.gpu_util import max_err, packed_doc_ids, rel_err, require_cuda
from touchnet_b200 import modeling

pytestmark = pytest.mark.gpu


class _Cfg:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def small_cfg(L=2, d=256, H=2, KV=1, ffn=512, V=512, rope_scaling=None, bias=False):
    return _Cfg(hidden_size=d, intermediate_size=ffn, num_hidden_layers=L, num_attention_heads=H, num_key_value_heads=KV,
                head_dim=128, vocab_size=V, rms_norm_eps=1e-5, rope_theta=500000.0, rope_scaling=rope_scaling,
                attention_bias=bias, tie_word_embeddings=False, initializer_range=0.02, model_type="llama",
                pad_token_id=0)


def oracle_cfg(c, audio=0):
    return mo.OracleConfig(hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
                           num_hidden_layers=c.num_hidden_layers, num_attention_heads=c.num_attention_heads,
                           num_key_value_heads=c.num_key_value_heads, head_dim=getattr(c, "head_dim", 128),
                           vocab_size=c.vocab_size,
                           rms_norm_eps=c.rms_norm_eps, rope_theta=c.rope_theta, rope_scaling=c.rope_scaling,
                           attention_bias=c.attention_bias, audio_input_size=audio,
                           pad_token_id=getattr(c, "pad_token_id", None))


LLAMA3 = {"rope_type": "llama3", "factor": 32.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
          "original_max_position_embeddings": 8192}


def _check_logits(ours, ref, valid):
    err = max_err(ours[valid].float(), ref[valid].float())
    scale = float(ref[valid].float().abs().max())
    assert err <= 3e-2 * scale + 1e-3, ("logits max err / scale", err, scale)
    top2 = ref[valid].float().topk(2, dim=-1).values
    decisive = (top2[:, 0] - top2[:, 1]) > 2 * err
    assert decisive.float().mean() > 0.3, ("fraction of decisive positions", float(decisive.float().mean()), err, scale)
    assert torch.equal(ours[valid].float().argmax(-1)[decisive], ref[valid].float().argmax(-1)[decisive])


@pytest.mark.parametrize("master_dtype", [torch.float32, torch.bfloat16])
def test_llama_forward_backward_parity(master_dtype):
    dev = require_cuda()
    cfg = small_cfg(rope_scaling=LLAMA3)
    B, T = 2, 384
    torch.manual_seed(2025)
    model = modeling.B200LlamaForCausalLM(cfg).to(dev)
    model.post_init()
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() == 2:
                p.normal_(0, 0.05)
            else:
                p.copy_(1 + 0.1 * torch.randn_like(p))
    model.to(master_dtype)
    doc, pos = packed_doc_ids(B, T, [[100, 200, 50], [384]], dev)
    ids = torch.randint(0, cfg.vocab_size, (B, T), device=dev)
    labels = torch.randint(0, cfg.vocab_size, (B, T), device=dev)
    labels[doc == 0] = -100
    out = model(input_ids=ids, attention_mask=doc, position_ids=pos)
    logits = out.logits
    assert logits.shape == (B, T, cfg.vocab_size) and logits.dtype == torch.bfloat16
    loss = torch.nn.functional.cross_entropy(logits.float().view(-1, cfg.vocab_size), labels.view(-1), ignore_index=-100)
    loss.backward()
    # oracle on the bf16-rounded weights: fp32 math (gradient reference) and bf16 math (rounding-point reference)
    params32 = {k: v.detach().bfloat16().float().requires_grad_(True) for k, v in model.state_dict().items()}
    ocfg = oracle_cfg(cfg)
    ref32 = mo.llama_forward(params32, ocfg, input_ids=ids, attention_mask=doc, position_ids=pos, dtype=torch.float32)
    with torch.no_grad():
        ref16 = mo.llama_forward({k: v.detach() for k, v in params32.items()}, ocfg, input_ids=ids, attention_mask=doc,
                                 position_ids=pos, dtype=torch.bfloat16)
    valid = doc > 0
    _check_logits(logits, ref32, valid)
    _check_logits(logits, ref16, valid)
    loss_ref = torch.nn.functional.cross_entropy(ref32.view(-1, cfg.vocab_size), labels.view(-1), ignore_index=-100)
    assert abs(float(loss) - float(loss_ref)) < 2e-2 * abs(float(loss_ref))
    loss_ref.backward()
    for name, p in model.named_parameters():
        assert p.grad is not None and p.grad.dtype == master_dtype, name
        assert torch.isfinite(p.grad.float()).all(), name
        e = rel_err(p.grad.float(), params32[name].grad)
        assert e < 3e-2, (name, e)


def test_touch_audio_forward_backward_parity():
    dev = require_cuda()
    F_ = 400                                                        # 80 mel x stack 5
    text = small_cfg(L=2, d=256, H=4, KV=4, bias=True)              # MHA + q/k/v bias: Qwen2-style text model (cfg 3)
    text.model_type = "qwen2"
    cfg = _Cfg(audio_config=_Cfg(input_size=F_), text_config=text, pad_token_id=0)
    B, T = 2, 256
    torch.manual_seed(7)
    model = modeling.B200TouchAudioForCausalLM(cfg).to(dev)
    model.post_init()
    with torch.no_grad():
        for n_, p in model.named_parameters():
            if p.dim() == 2:
                p.normal_(0, 0.05)
            elif "bias" in n_:
                p.normal_(0, 0.1)
    doc, pos = packed_doc_ids(B, T, [[120, 100], [200, 56]], dev)
    is_audio = torch.zeros(B, T, dtype=torch.bool, device=dev)
    is_audio[0, :90] = True; is_audio[0, 120:200] = True; is_audio[1, :150] = True; is_audio[1, 200:240] = True
    feats = torch.randn(B, T, F_, device=dev) * is_audio[..., None]
    ids = torch.where(is_audio, torch.zeros(B, T, dtype=torch.int64, device=dev),
                      torch.randint(3, text.vocab_size, (B, T), device=dev))
    out = model(input_ids=ids, input_features=feats, attention_mask=doc, position_ids=pos, inputs_embeds=None)
    assert out.attention_mask is doc                                # ref: modeling_touch_audio.py:151
    logits = out.logits
    logits.float().square().mean().backward()
    model.raise_if_nan()
    params32 = {k: v.detach().bfloat16().float().requires_grad_(True) for k, v in model.state_dict().items()}
    ref = mo.touch_audio_forward(params32, oracle_cfg(text, audio=F_), input_ids=ids,
                                 input_features=feats.bfloat16().float(), attention_mask=doc, position_ids=pos)
    _check_logits(logits, ref, doc > 0)
    # padding rows feed zeros forward in ours (FlexAttention semantics); restrict the loss to valid rows for gradients
    model.zero_grad()
    out = model(input_ids=ids, input_features=feats, attention_mask=doc, position_ids=pos)
    (out.logits.float()[doc > 0]).square().mean().backward()
    ref[doc > 0].square().mean().backward()
    for name, p in model.named_parameters():
        e = rel_err(p.grad.float(), params32[name].grad)
        assert e < 3e-2, (name, e)
    # NaN in the features -> ValueError("NaN in data.") raised by forward itself, as ref modeling_touch_audio.py:133-134
    feats[0, 3, 5] = float("nan")
    with pytest.raises(ValueError, match="NaN in data"):
        model(input_ids=ids, input_features=feats, attention_mask=doc, position_ids=pos)
    model(input_ids=ids, input_features=feats, attention_mask=doc, position_ids=pos, check_nan=False)   # deferred form
    with pytest.raises(ValueError, match="NaN in data"):
        model.raise_if_nan()
    feats[0, 3, 5] = 0.0
    model(input_ids=ids, input_features=feats, attention_mask=doc, position_ids=pos)    # the flag was cleared


@pytest.mark.parametrize("hd,H,KV", [(64, 4, 2), (32, 8, 8)])
def test_head_dim_below_128_runs_through_zero_padded_heads(hd, H, KV):
    """The reference's example config is head_dim 64 (examples/text/pretrain/fineweb-edu/config/Llama-3_2-1B.json:13): such
    models run through the 128-wide attention kernels with zero-padded heads (exact: zero q/k columns add nothing to the
    scores, zero v columns give zero outputs).  Forward logits and every parameter gradient vs the oracle."""
    dev = require_cuda()
    cfg = small_cfg(L=2, d=256, H=H, KV=KV, ffn=512, V=512, rope_scaling=LLAMA3)
    cfg.head_dim = hd
    B, T = 2, 384
    torch.manual_seed(2025)
    model = modeling.B200LlamaForCausalLM(cfg).to(dev)
    model.post_init()
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() == 2:
                p.normal_(0, 0.05)
    doc, pos = packed_doc_ids(B, T, [[100, 200, 50], [384]], dev)
    ids = torch.randint(1, cfg.vocab_size, (B, T), device=dev)
    logits = model(input_ids=ids, attention_mask=doc, position_ids=pos).logits
    (logits.float()[doc > 0]).square().mean().backward()
    params32 = {k: v.detach().bfloat16().float().requires_grad_(True) for k, v in model.state_dict().items()}
    ref = mo.llama_forward(params32, oracle_cfg(cfg), input_ids=ids, attention_mask=doc, position_ids=pos, dtype=torch.float32)
    _check_logits(logits, ref, doc > 0)
    ref[doc > 0].square().mean().backward()
    for name, p in model.named_parameters():
        e = rel_err(p.grad.float(), params32[name].grad)
        assert e < 3e-2, (name, e)


@pytest.mark.parametrize("mode", ["full", "selective_op"])
def test_activation_checkpointing_wrappers_are_exact(mode):
    """SURVEY 2 row 6 (compatibility constraint): the reference wraps every decoder block in `checkpoint_wrapper`
    (full: touchnet/models/helper_func.py:48-49; op-selective SAC: :58-85, run.sh:158 uses full).  The fused block must
    recompute to the same bits: loss and every gradient equal the un-checkpointed run exactly (deterministic kernels;
    the in-place RoPE / residual epilogues live inside one autograd node, so recomputation sees pristine inputs)."""
    from collections import defaultdict
    from torch.distributed.algorithms._checkpoint.checkpoint_wrapper import checkpoint_wrapper
    dev = require_cuda()
    cfg = small_cfg(L=3, rope_scaling=LLAMA3)
    B, T = 2, 384
    doc, pos = packed_doc_ids(B, T, [[100, 200, 50], [384]], dev)
    torch.manual_seed(11)
    ids = torch.randint(1, cfg.vocab_size, (B, T), device=dev)
    labels = torch.randint(0, cfg.vocab_size, (B, T), device=dev)
    labels[doc == 0] = -100

    def build():
        torch.manual_seed(2025)
        m = modeling.B200LlamaForCausalLM(cfg).to(dev)
        m.post_init()
        with torch.no_grad():
            for p in m.parameters():
                if p.dim() == 2:
                    p.normal_(0, 0.05)
        return m

    def run(m):
        logits = m(input_ids=ids, attention_mask=doc, position_ids=pos).logits
        loss = torch.nn.functional.cross_entropy(logits.float().view(-1, cfg.vocab_size), labels.view(-1), ignore_index=-100)
        loss.backward()
        return loss.detach(), {n.replace("._checkpoint_wrapped_module", ""): p.grad.clone() for n, p in m.named_parameters()}

    loss0, g0 = run(build())
    m = build()
    layers = m.model.layers
    for name, block in list(layers.named_children()):
        if mode == "full":
            wrapped = checkpoint_wrapper(block, preserve_rng_state=False)
        else:
            from torch.utils.checkpoint import CheckpointPolicy, create_selective_checkpoint_contexts

            def ctx_fn():
                meta = defaultdict(int)

                def policy(ctx, func, *args, **kwargs):        # the reference's policy: save every 2nd aten.mm
                    if func == torch.ops.aten.mm.default:
                        meta["mm"] += 1
                        return CheckpointPolicy.MUST_SAVE if meta["mm"] % 2 else CheckpointPolicy.PREFER_RECOMPUTE
                    return CheckpointPolicy.PREFER_RECOMPUTE
                return create_selective_checkpoint_contexts(policy)
            wrapped = checkpoint_wrapper(block, context_fn=ctx_fn, preserve_rng_state=False)
        layers.register_module(name, wrapped)
    loss1, g1 = run(m)
    assert torch.equal(loss0, loss1)
    assert g0.keys() == g1.keys()
    for n in g0:
        assert torch.equal(g0[n], g1[n]), n


def test_plain_causal_defaults_and_text_only_batches():
    dev = require_cuda()
    cfg = small_cfg(L=1)
    torch.manual_seed(3)
    model = modeling.B200LlamaForCausalLM(cfg).to(dev)
    model.post_init()
    ids = torch.randint(0, cfg.vocab_size, (1, 200), device=dev)    # T not a multiple of 128, no mask, no positions
    logits = model(input_ids=ids).logits
    params = {k: v.detach().bfloat16().float() for k, v in model.state_dict().items()}
    ref = mo.llama_forward(params, oracle_cfg(cfg), input_ids=ids)
    _check_logits(logits, ref, torch.ones(1, 200, dtype=torch.bool, device=dev))


def test_matches_hf_flex_attention_path():
    """The thing the reference actually runs on a GPU: HF LlamaForCausalLM + flex_attention BlockMask
    (make_flex_block_causal_mask).  Skipped only if this transformers/torch build cannot construct it."""
    dev = require_cuda()
    try:
        from transformers import LlamaConfig, LlamaForCausalLM
        from transformers.integrations.flex_attention import make_flex_block_causal_mask
    except Exception as e:  # pragma: no cover
        pytest.skip(f"transformers flex integration unavailable: {e}")
    cfg = small_cfg(L=2, rope_scaling=LLAMA3)
    hf_cfg = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                         num_key_value_heads=1, head_dim=128, vocab_size=512, rms_norm_eps=1e-5, rope_theta=500000.0,
                         rope_scaling=dict(LLAMA3), tie_word_embeddings=False, attention_bias=False)
    hf_cfg._attn_implementation = "flex_attention"
    torch.manual_seed(11)
    try:
        hf = LlamaForCausalLM(hf_cfg)
        with torch.no_grad():
            for p_ in hf.parameters():
                if p_.dim() == 2:
                    p_.normal_(0, 0.05)       # logits of O(1) so that the comparison is meaningful
        hf = hf.to(dev).to(torch.bfloat16).eval()
        B, T = 1, 256
        doc, pos = packed_doc_ids(B, T, [[100, 120]], dev)
        ids = torch.randint(0, 512, (B, T), device=dev)
        bm = make_flex_block_causal_mask(doc)
        with torch.no_grad():
            ref = hf(input_ids=ids, attention_mask=bm, position_ids=pos).logits
    except Exception as e:
        pytest.skip(f"HF flex_attention path not runnable in this image: {type(e).__name__}: {str(e)[:200]}")
    ours = modeling.B200LlamaForCausalLM(cfg).to(dev)
    missing = ours.load_state_dict(hf.state_dict(), strict=True)     # identical state-dict keys
    # (HF recomputes the rope frequencies in fp32 even when the module was cast to bf16; ours does the same)
    logits = ours(input_ids=ids, attention_mask=doc, position_ids=pos).logits
    _check_logits(logits, ref, doc > 0)


def test_stock_hf_llama_with_b200_kernels_patched_in():
    """The inner seam: stock HF LlamaForCausalLM whose RMSNorm / MLP / RoPE / attention arithmetic is
    swapped Liger-style for the B200 kernels == the same HF model run eagerly with the dense document mask,
    forward and parameter gradients (both bf16 models, so tolerances are bf16-vs-bf16)."""
    dev = require_cuda()
    import importlib
    from transformers import LlamaConfig, LlamaForCausalLM
    from transformers.models.llama import modeling_llama
    from touchnet_b200 import train_spec
    hf_cfg = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                         num_key_value_heads=1, head_dim=128, vocab_size=512, rms_norm_eps=1e-5, rope_theta=500000.0,
                         rope_scaling=dict(LLAMA3), tie_word_embeddings=False, attention_bias=False)
    hf_cfg._attn_implementation = "eager"
    torch.manual_seed(5)
    hf = LlamaForCausalLM(hf_cfg)
    with torch.no_grad():
        for p_ in hf.parameters():
            if p_.dim() == 2:
                p_.normal_(0, 0.05)
    hf = hf.to(dev).to(torch.bfloat16)
    B, T = 2, 256
    doc, pos = packed_doc_ids(B, T, [[100, 120], [256]], dev)
    ids = torch.randint(0, 512, (B, T), device=dev)
    allow = mo.doc_causal_allow(doc)[:, None]                          # [B,1,T,T] bool
    mask4 = torch.zeros(allow.shape, dtype=torch.bfloat16, device=dev).masked_fill(~allow, float("-inf"))
    mask4[(doc == 0)[:, None, :, None].expand_as(mask4)] = 0.0         # keep padding rows finite in eager softmax
    tgt = torch.randn(B, T, 512, device=dev)
    valid = doc > 0

    def run(**kw):
        hf.zero_grad()
        lg = hf(input_ids=ids, position_ids=pos, **kw).logits
        ((lg.float() * tgt)[valid]).mean().backward()
        return lg.detach(), {n: p.grad.detach().clone() for n, p in hf.named_parameters()}

    ref, ref_g = run(attention_mask=mask4)
    saved = (modeling_llama.apply_rotary_pos_emb, modeling_llama.LlamaRMSNorm.forward, modeling_llama.LlamaMLP.forward)
    try:
        train_spec.apply_b200_kernels_to_hf_llama()
        hf.config._attn_implementation = "touchnet_b200"
        with train_spec.packed_document_ids(doc):
            ours, ours_g = run(attention_mask=None)
    finally:
        (modeling_llama.apply_rotary_pos_emb, modeling_llama.LlamaRMSNorm.forward, modeling_llama.LlamaMLP.forward) = saved
        hf.config._attn_implementation = "eager"
    _check_logits(ours, ref, valid)
    for n in ref_g:
        e = rel_err(ours_g[n].float(), ref_g[n].float())
        assert e < 4e-2, (n, e)
