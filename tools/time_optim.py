"""Time tn_adamw_f32 / tn_sumsq_f32 on one Llama-3-8B block's worth of parameters (218.1 M fp32) against HBM bandwidth."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from touchnet_b200 import optim, ops
dev = torch.device("cuda")
n = 218_112_000
p = torch.nn.Parameter(torch.randn(n // 4096, 4096, device=dev))
p.grad = torch.randn_like(p)
ops.bf16_weight(p)                       # creates the working copy the step keeps fresh
o = optim.B200AdamW([p], lr=1e-4)
ref_p = torch.nn.Parameter(p.detach().clone()); ref_p.grad = p.grad.clone()
o_ref = torch.optim.AdamW([ref_p], lr=1e-4, betas=(0.9, 0.95), weight_decay=0.1, fused=True)
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
t_ours = timeit(lambda: (optim.clip_grad_norm_([p], 1.0, defer_to=o), o.step()))
t_ref = timeit(lambda: (torch.nn.utils.clip_grad_norm_([ref_p], 1.0), o_ref.step(), ref_p.detach().bfloat16()))
bytes_ours = n * (4 + 16 + 12 + 2)      # norm pass reads g; step reads p,g,m,v, writes p,m,v + bf16 copy
print(f"clip+AdamW+bf16 copy, {n/1e6:.1f} M params: ours {t_ours*1e3:.0f} us ({bytes_ours/t_ours/1e6:.0f} GB/s algorithmic), "
      f"torch clip_grad_norm_ + fused AdamW + cast {t_ref*1e3:.0f} us -> x{t_ref/t_ours:.2f}")
