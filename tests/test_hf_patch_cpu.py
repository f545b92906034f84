"""Host logic of the inner seam on CPU: `apply_b200_kernels_to_hf_llama()` + the "touchnet_b200" attention
implementation + `packed_document_ids` plumb the stock HF LlamaForCausalLM correctly (mask registry, RoPE layout
conversion, autograd through every swapped op).  The CUDA ops are replaced by the plain-torch stand-ins of
tests/cpu_ops_shim.py, so this checks the wiring only; the same comparison through the real kernels is
tests/test_gpu_model.py::test_stock_hf_llama_with_b200_kernels_patched_in."""
import pytest
import torch

from oracle import model_oracle as mo
from tests import cpu_ops_shim
from tests.gpu_util import packed_doc_ids, rel_err


def test_patched_hf_llama_matches_hf_eager_on_cpu():
    transformers = pytest.importorskip("transformers")
    from transformers import LlamaConfig, LlamaForCausalLM
    from transformers.models.llama import modeling_llama
    from touchnet_b200 import train_spec
    cfg = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                      num_key_value_heads=1, head_dim=128, vocab_size=512, rms_norm_eps=1e-5, rope_theta=500000.0,
                      tie_word_embeddings=False, attention_bias=False)
    cfg._attn_implementation = "eager"
    torch.manual_seed(5)
    hf = LlamaForCausalLM(cfg)
    with torch.no_grad():
        for p in hf.parameters():
            if p.dim() == 2:
                p.normal_(0, 0.05)
    hf = hf.to(torch.bfloat16)
    B, T = 2, 256
    doc, pos = packed_doc_ids(B, T, [[100, 120], [256]])
    ids = torch.randint(0, 512, (B, T))
    allow = mo.doc_causal_allow(doc)[:, None]
    mask4 = torch.zeros(allow.shape, dtype=torch.bfloat16).masked_fill(~allow, float("-inf"))
    mask4[(doc == 0)[:, None, :, None].expand_as(mask4)] = 0.0
    tgt = torch.randn(B, T, 512)
    valid = doc > 0

    def run(**kw):
        hf.zero_grad()
        lg = hf(input_ids=ids, position_ids=pos, **kw).logits
        ((lg.float() * tgt)[valid]).mean().backward()
        return lg.detach(), {n: p.grad.detach().clone() for n, p in hf.named_parameters()}

    ref, ref_g = run(attention_mask=mask4)
    saved_hf = (modeling_llama.apply_rotary_pos_emb, modeling_llama.LlamaRMSNorm.forward, modeling_llama.LlamaMLP.forward)
    saved_ops = cpu_ops_shim.install()
    try:
        train_spec.apply_b200_kernels_to_hf_llama()
        hf.config._attn_implementation = "touchnet_b200"
        with train_spec.packed_document_ids(doc):
            ours, ours_g = run(attention_mask=None)
    finally:
        cpu_ops_shim.uninstall(saved_ops)
        (modeling_llama.apply_rotary_pos_emb, modeling_llama.LlamaRMSNorm.forward, modeling_llama.LlamaMLP.forward) = saved_hf
    scale = float(ref.float().abs().max())
    assert float((ours.float() - ref.float())[valid].abs().max()) < 3e-2 * scale
    for n in ref_g:
        assert rel_err(ours_g[n].float(), ref_g[n].float()) < 4e-2, n
