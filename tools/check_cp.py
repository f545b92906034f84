"""2-rank (or cp-rank) check of context parallelism: CP-sharded forward/backward == the unsharded run.
torchrun --nproc-per-node 2 tools/check_cp.py"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from touchnet_b200 import modeling, context_parallel
from tests.gpu_util import packed_doc_ids, rel_err
from tests.test_gpu_model import small_cfg

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
cfg = small_cfg(L=2, d=512, H=4, KV=2, ffn=1024, V=512)
torch.manual_seed(1)
model = modeling.B200LlamaForCausalLM(cfg).to(dev); model.post_init()
with torch.no_grad():
    for p in model.parameters():
        if p.dim() == 2: p.normal_(0, 0.05)
for p in model.parameters():
    dist.broadcast(p.data, 0)
B, T = 2, 1024
doc, pos = packed_doc_ids(B, T, [[300, 500, 100], [1024]], dev)
g = torch.Generator(device="cpu").manual_seed(3)
ids = torch.randint(0, cfg.vocab_size, (B, T), generator=g).to(dev)
tgt = torch.randn(B, T, cfg.vocab_size, generator=g).to(dev)
# ---- unsharded reference (every rank computes it) ----
logits = model(input_ids=ids, attention_mask=doc, position_ids=pos).logits
((logits.float() * tgt)[doc > 0]).mean().backward()
ref_logits = logits.detach().clone()
ref_grads = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
model.zero_grad()
n_valid = int((doc > 0).sum())
# ---- context parallel: every rank holds T/world rows ----
Tl = T // world
sl = slice(rank * Tl, (rank + 1) * Tl)
context_parallel.enable_context_parallel(model, dist.group.WORLD)
lg = model(input_ids=ids[:, sl].contiguous(), attention_mask=doc[:, sl].contiguous(), position_ids=pos[:, sl].contiguous()).logits
loss = ((lg.float() * tgt[:, sl])[doc[:, sl] > 0]).sum() / (n_valid * cfg.vocab_size)   # same global mean
loss.backward()
err_fwd = float((lg.float() - ref_logits[:, sl].float()).abs().max())
worst = 0.0
for n, p in model.named_parameters():
    gsum = p.grad.detach().clone()
    dist.all_reduce(gsum)                                                    # ranks hold partial sums over their tokens
    e = rel_err(gsum.float(), ref_grads[n].float())
    if rank == 0 and (e > 2e-2 or os.environ.get("CP_VERBOSE")):
        print(f"  {n}: rel err {e:.4g} |ref| {float(ref_grads[n].float().norm()):.4g} |cp| {float(gsum.float().norm()):.4g}")
    worst = max(worst, e)
res = torch.tensor([err_fwd, worst], device=dev)
dist.all_reduce(res, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"CP{world}: max |logits - unsharded| = {res[0].item():.4g} (scale {float(ref_logits.float().abs().max()):.3g}); worst grad rel err = {res[1].item():.4g}")
    assert res[0].item() < 2e-2 * float(ref_logits.float().abs().max()) and res[1].item() < 2e-2
    print("CP OK")
dist.destroy_process_group()
