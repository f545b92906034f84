"""2-rank check of FSDP2 over the B200 modules: loss and (all-gathered) gradients equal the single-process run on the
concatenated batch.  torchrun --nproc-per-node 2 tools/check_fsdp.py"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch.distributed.fsdp import fully_shard, MixedPrecisionPolicy
from torch.distributed.device_mesh import init_device_mesh
from touchnet_b200 import modeling
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
T = 1024
lens = [[300, 500, 100], [1024], [512, 512], [1000]][: world]
doc, pos = packed_doc_ids(world, T, lens, dev)
g = torch.Generator(device="cpu").manual_seed(3)
ids = torch.randint(0, cfg.vocab_size, (world, T), generator=g).to(dev)
labels = torch.randint(0, cfg.vocab_size, (world, T), generator=g).to(dev)
labels[doc == 0] = -100


def step(m, sl):
    lg = m(input_ids=ids[sl], attention_mask=doc[sl], position_ids=pos[sl]).logits
    ce = torch.nn.functional.cross_entropy(lg.float().view(-1, cfg.vocab_size), labels[sl].reshape(-1), ignore_index=-100,
                                           reduction="sum")
    return ce / int((labels != -100).sum())                      # global token mean, so rank losses add up


# ---- single-process reference over the whole batch ----
ref_loss = step(model, slice(0, world))
ref_loss.backward()
ref_grads = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
LR = 20.0
with torch.no_grad():                                            # one SGD step, then the loss again (second-step check)
    for p in model.parameters():
        p.add_(p.grad, alpha=-LR)
ref_loss2 = step(model, slice(0, world)).detach()
with torch.no_grad():
    for p in model.parameters():
        p.add_(p.grad, alpha=LR)                                 # undo
model.zero_grad()
# ---- FSDP2: one row per rank ----
mesh = init_device_mesh("cuda", (world,), mesh_dim_names=("dp_shard",))
fp32_params = os.environ.get("FSDP_FP32", "0") == "1"           # param_dtype=None: fp32 all-gather, the layer casts itself
mp = MixedPrecisionPolicy(param_dtype=None if fp32_params else torch.bfloat16, reduce_dtype=torch.float32)
for layer in model.model.layers:
    fully_shard(layer, mesh=mesh, mp_policy=mp)
fully_shard(model, mesh=mesh, mp_policy=mp)
if os.environ.get("TN_FSDP_PEER", "0") != "0":                   # our pull kernels over NVLink peer memory instead of NCCL
    from touchnet_b200 import fsdp_comm
    fsdp_comm.install(model, mesh.get_group(), dev, max_ctas=int(os.environ.get("TN_FSDP_PEER_CTAS", "32")),
                              mode="push" if os.environ["TN_FSDP_PEER"] == "push" else "pull",
                              direct=os.environ.get("TN_FSDP_DIRECT", "1") != "0")
    if rank == 0:
        print("FSDP2 collectives: tn_peer_reduce_scatter_f32 / tn_peer_all_gather over symmetric memory")
loss = step(model, slice(rank, rank + 1))
loss.backward()
tot = loss.detach().clone()
dist.all_reduce(tot)
worst = 0.0
for n, p in model.named_parameters():
    gfull = p.grad.full_tensor() * world                         # FSDP averages over ranks; the rank losses are partial sums
    e = rel_err(gfull.float(), ref_grads[n].float())
    if rank == 0 and e > 2e-2:
        print(f"  {n}: rel err {e:.4g}")
    worst = max(worst, e)
with torch.no_grad():
    for p in model.parameters():
        p.add_(p.grad * world, alpha=-LR)
loss2 = step(model, slice(rank, rank + 1)).detach()
dist.all_reduce(loss2)
if rank == 0:
    print(f"second step: loss {loss2.item():.6f} vs {ref_loss2.item():.6f} (first {ref_loss.item():.6f})")
    assert abs(ref_loss2.item() - ref_loss.item()) > 1e-2 * abs(ref_loss.item()), "update too small to test anything"
    assert abs(loss2.item() - ref_loss2.item()) < 3e-3 * abs(ref_loss2.item())
if rank == 0:
    print(f"FSDP{world}: loss {tot.item():.6f} vs {ref_loss.item():.6f}; worst grad rel err {worst:.4g}")
    assert abs(tot.item() - ref_loss.item()) < 2e-3 * abs(ref_loss.item()) and worst < 2e-2
    print("FSDP OK")
dist.destroy_process_group()
