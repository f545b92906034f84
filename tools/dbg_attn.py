import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model_oracle as mo
from tests.gpu_util import packed_doc_ids, rel_err
from touchnet_b200 import ops, batching
dev = torch.device("cuda:0")
sc = 1 / math.sqrt(128)
def run(name, B, T, H, KV, doc):
    g = torch.Generator(device="cpu").manual_seed(1)
    mk = lambda c: torch.randn(B * T, c * 128, generator=g).to(dev).bfloat16()
    q, k, v, do = mk(H), mk(KV), mk(KV), mk(H)
    plan = ops.AttnPlan(doc)
    o, lse = ops.attn_fwd(q, k, v, plan, H, KV, sc)
    o2, lse2 = ops.attn_fwd(q, k, v, plan, H, KV, sc)
    dq, dk, dv = ops.attn_bwd(q, k, v, o, do, lse, plan, H, KV, sc)
    dq2, dk2, dv2 = ops.attn_bwd(q, k, v, o, do, lse, plan, H, KV, sc)
    torch.backends.cuda.matmul.allow_tf32 = False
    o_r, lse_r, dq_r, dk_r, dv_r = mo.attention_chunked(q, k, v, doc, H, KV, sc, do, q_chunk=4096)
    print(name, "det fwd", torch.equal(o, o2), "det bwd", torch.equal(dq, dq2), torch.equal(dk, dk2), torch.equal(dv, dv2))
    for nm, a, r in (("o", o, o_r), ("dq", dq, dq_r), ("dk", dk, dk_r), ("dv", dv, dv_r)):
        e = rel_err(a.float(), r)
        # per 128-row block error to localise
        d = (a.float() - r).view(B, T // 128, 128, -1).pow(2).sum((2, 3)).sqrt() / (r.view(B, T // 128, 128, -1).pow(2).sum((2, 3)).sqrt() + 1e-20)
        bad = (d > 2e-2).nonzero().tolist()
        print("  ", nm, "rel", round(e, 5), "bad blocks", bad[:12], len(bad))
doc, _ = packed_doc_ids(1, 1024, [[300, 500, 100]], dev)
run("T1024 3docs H4", 1, 1024, 4, 2, doc)
doc = batching.plan_audio_text_batch(2025, 1, 8192, 128256, stride=4, max_s=30.0)[0]["attention_mask"].to(dev)
run("cfg2", 1, 8192, 32, 8, doc)
print("doc starts", [int(x) for x in (doc[0][1:] != doc[0][:-1]).nonzero().flatten()[:25] + 1])
