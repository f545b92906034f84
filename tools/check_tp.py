"""2-rank check of tensor + sequence parallelism on the real kernels: TP-sharded forward/backward == the unsharded run.
torchrun --nproc-per-node 2 tools/check_tp.py      (TP_AUDIO=1: TouchAudio model with q/k/v bias)"""
import copy, os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch.distributed.device_mesh import init_device_mesh
from touchnet_b200 import modeling, tensor_parallel
from tests.gpu_util import packed_doc_ids, rel_err
from tests.test_gpu_model import small_cfg, _Cfg

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
audio = os.environ.get("TP_AUDIO", "0") == "1"
text = small_cfg(L=2, d=512, H=4, KV=2, ffn=1024, V=512, bias=audio)
torch.manual_seed(1)
if audio:
    model = modeling.B200TouchAudioForCausalLM(_Cfg(audio_config=_Cfg(input_size=400), text_config=text, pad_token_id=0)).to(dev)
else:
    model = modeling.B200LlamaForCausalLM(text).to(dev)
model.post_init()
with torch.no_grad():
    for n, p in model.named_parameters():
        if p.dim() == 2: p.normal_(0, 0.05)
        elif "bias" in n: p.normal_(0, 0.1)
        else: p.uniform_(0.5, 1.5)
for p in model.parameters():
    dist.broadcast(p.data, 0)
B, T = 2, 1024
doc, pos = packed_doc_ids(B, T, [[300, 500, 100], [1024]], dev)
g = torch.Generator(device="cpu").manual_seed(3)
ids = torch.randint(1, text.vocab_size, (B, T), generator=g).to(dev)
tgt = torch.randn(B, T, text.vocab_size, generator=g).to(dev)
kw = dict(input_ids=ids, attention_mask=doc, position_ids=pos)
if audio:
    is_audio = torch.zeros(B, T, dtype=torch.bool, device=dev); is_audio[:, :256] = True
    kw["input_features"] = torch.randn(B, T, 400, generator=g).to(dev) * is_audio[..., None]
    kw["input_ids"] = torch.where(is_audio, torch.zeros_like(ids), ids)
ref_model = copy.deepcopy(model)
ref_logits = ref_model(**kw).logits
((ref_logits.float() * tgt)[doc > 0]).mean().backward()
ref_grads = {n: p.grad.detach().clone() for n, p in ref_model.named_parameters()}
mesh = init_device_mesh("cuda", (world,), mesh_dim_names=("tp",))
tensor_parallel.apply_tp(model, mesh)
logits = model(**kw).logits
((logits.float() * tgt)[doc > 0]).mean().backward()
scale = float(ref_logits.float().abs().max())
err_fwd = float((logits.float() - ref_logits.float())[doc > 0].abs().max())
worst = 0.0
for n, p in model.named_parameters():
    gfull = p.grad.full_tensor() if tensor_parallel.is_dtensor(p.grad) else p.grad
    e = rel_err(gfull.float(), ref_grads[n].float())
    if rank == 0 and (e > 3e-2 or os.environ.get("TP_VERBOSE")):
        print(f"  {n}: rel err {e:.4g}")
    worst = max(worst, e)
res = torch.tensor([err_fwd, worst], device=dev)
dist.all_reduce(res, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"TP{world}{' audio' if audio else ''}: max |logits - unsharded| = {res[0].item():.4g} (scale {scale:.3g}); worst grad rel err = {res[1].item():.4g}")
    assert res[0].item() < 3e-2 * scale and res[1].item() < 3e-2
    print("TP OK")
# ---- loss parallel (ref: parallelize_llama.py:127-131 + distributed.py:322-323): vocabulary-sharded logits into the pack-loss ----
if os.environ.get("TP_LOSS_PARALLEL", "1") != "0":
    from touchnet_b200 import loss as tn_loss
    labels = torch.randint(0, text.vocab_size, (B, T), generator=g).to(dev)
    labels[doc == 0] = -100
    sl = torch.randint(1, 30, (B, T), generator=g).to(dev)
    ref_model.zero_grad(); model.zero_grad()
    lps_ref, lpt_ref = tn_loss.cross_entropy_loss(ref_model(**kw).logits, labels, sl, 7)
    acc_ref = tn_loss.accuracy(ref_model(**kw).logits, labels)
    lps_ref.backward()
    ref_grads = {n: p.grad.detach().clone() for n, p in ref_model.named_parameters()}
    lm = model.language_model if hasattr(model, "language_model") else model
    lm.loss_parallel = True
    model.train()
    pred = model(**kw).logits
    assert isinstance(pred, tn_loss.VocabParallelLogits) and pred.local.shape[-1] == text.vocab_size // world
    lps, lpt = tn_loss.cross_entropy_loss(pred, labels, sl, 7)
    acc = tn_loss.accuracy(pred, labels)
    lps.backward()
    worst = 0.0
    for n, p in model.named_parameters():
        gfull = p.grad.full_tensor() if tensor_parallel.is_dtensor(p.grad) else p.grad
        worst = max(worst, rel_err(gfull.float(), ref_grads[n].float()))
    res = torch.tensor([abs(float(lps) - float(lps_ref)) / abs(float(lps_ref)), abs(float(lpt) - float(lpt_ref)) / abs(float(lpt_ref)),
                        abs(float(acc) - float(acc_ref)), worst], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"TP{world} loss parallel: rel loss err {res[0].item():.3g} / {res[1].item():.3g}, |acc diff| {res[2].item():.3g}, "
              f"worst grad rel err {res[3].item():.4g}")
        assert res[0].item() < 2e-3 and res[1].item() < 2e-3 and res[2].item() < 2e-3 and res[3].item() < 3e-2
        print("TP LOSS PARALLEL OK")
dist.destroy_process_group()
