"""Diagnose the HF-flex model-level comparison: which of {HF flex, HF eager + 4-D mask, oracle, ours} disagree."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformers import LlamaConfig, LlamaForCausalLM
from transformers.integrations.flex_attention import make_flex_block_causal_mask
from oracle import model_oracle as mo
from touchnet_b200 import modeling
from tests.gpu_util import packed_doc_ids
from tests.test_gpu_model import small_cfg, oracle_cfg, LLAMA3
dev = torch.device("cuda")
def mk(impl):
    c = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=1,
                    head_dim=128, vocab_size=512, rms_norm_eps=1e-5, rope_theta=500000.0, rope_scaling=dict(LLAMA3),
                    tie_word_embeddings=False, attention_bias=False)
    c._attn_implementation = impl
    return c
torch.manual_seed(11)
hf = LlamaForCausalLM(mk("flex_attention"))
with torch.no_grad():
    for p_ in hf.parameters():
        if p_.dim() == 2: p_.normal_(0, 0.05)
sd = {k: v.clone() for k, v in hf.state_dict().items()}
print("hf config rope:", getattr(hf.config, "rope_scaling", None), getattr(hf.config, "rope_parameters", None), getattr(hf.config, "rope_theta", None))
print("hf inv_freq[:4] fp32:", hf.model.rotary_emb.inv_freq[:4].tolist(), "attn_scaling", hf.model.rotary_emb.attention_scaling)
B, T = 1, 256
doc, pos = packed_doc_ids(B, T, [[100, 120]], dev)
ids = torch.randint(0, 512, (B, T), device=dev)
idx = torch.arange(T, device=dev)
allow = (idx[:, None] >= idx[None, :])[None] & (doc[:, :, None] == doc[:, None, :]) & (doc > 0)[:, :, None]
res = {}
for dt in (torch.float32, torch.bfloat16):
    hf_f = LlamaForCausalLM(mk("flex_attention")); hf_f.load_state_dict(sd); hf_f = hf_f.to(dev).to(dt).eval()
    inv32 = LlamaForCausalLM(mk("eager")).model.rotary_emb.inv_freq.clone()
    with torch.no_grad():
        try:
            res[("flex", dt)] = hf_f(input_ids=ids, attention_mask=make_flex_block_causal_mask(doc), position_ids=pos).logits.float()
        except Exception as e:
            print("flex failed", dt, repr(e)[:300])
    hf_e = LlamaForCausalLM(mk("eager")); hf_e.load_state_dict(sd); hf_e = hf_e.to(dev).to(dt).eval()
    m4 = torch.zeros(B, 1, T, T, device=dev, dtype=dt).masked_fill(~allow[:, None], torch.finfo(dt).min)
    with torch.no_grad():
        res[("eager", dt)] = hf_e(input_ids=ids, attention_mask=m4, position_ids=pos).logits.float()
    print(dt, "hf inv_freq dtype after .to:", hf_e.model.rotary_emb.inv_freq.dtype, hf_e.model.rotary_emb.inv_freq[:3].tolist())
params = {k: v.to(dev).bfloat16().float() for k, v in sd.items()}
cfg = small_cfg(L=2, rope_scaling=LLAMA3)
res["oracle32"] = mo.llama_forward(params, oracle_cfg(cfg), input_ids=ids, attention_mask=doc, position_ids=pos)
ours = modeling.B200LlamaForCausalLM(cfg).to(dev); ours.load_state_dict({k: v.to(dev) for k, v in sd.items()})
res["ours"] = ours(input_ids=ids, attention_mask=doc, position_ids=pos).logits.float()
v = (doc > 0)
keys = list(res)
for i, a in enumerate(keys):
    for b in keys[i + 1:]:
        print(f"{str(a):38s} vs {str(b):38s} max err {float((res[a][v] - res[b][v]).abs().max()):.4f}  (scale {float(res[a][v].abs().max()):.3f})")
