"""Stand-alone timing of the HBM-bound kernels at the bench shape ([8192, 4096] rows, ffn 14336), rotating over more
buffers than fit in L2.  Prints achieved GB/s on algorithmic bytes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from touchnet_b200 import ops
dev = torch.device("cuda")
rows, d, ffn, NB = 8192, 4096, 14336, 6
xs = [torch.randn(rows, d, device=dev).bfloat16() for _ in range(NB)]
dys = [torch.randn(rows, d, device=dev).bfloat16() for _ in range(NB)]
ex = [torch.randn(rows, d, device=dev).bfloat16() for _ in range(NB)]
w = torch.ones(d, device=dev)
def timeit(fn, iters=24):
    for i in range(6): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
rstd = ops.rmsnorm_fwd(xs[0], w, 1e-5)[2]
t = timeit(lambda i: ops.rmsnorm_fwd(xs[i % NB], w, 1e-5))
print(f"rmsnorm_fwd  {t:7.1f} us  {rows*d*4/t/1e3:7.0f} GB/s")
t = timeit(lambda i: ops.rmsnorm_bwd(xs[i % NB], dys[i % NB], w, rstd, ds_extra=ex[i % NB]))
print(f"rmsnorm_bwd  {t:7.1f} us  {rows*d*8/t/1e3:7.0f} GB/s (incl. the column-sum kernel)")
g = [torch.randn(rows, ffn, device=dev).bfloat16() for _ in range(3)]
u = [torch.randn(rows, ffn, device=dev).bfloat16() for _ in range(3)]
dh = [torch.randn(rows, ffn, device=dev).bfloat16() for _ in range(3)]
t = timeit(lambda i: ops.swiglu_bwd(g[i % 3], u[i % 3], dh[i % 3]))
print(f"swiglu_bwd   {t:7.1f} us  {rows*ffn*10/t/1e3:7.0f} GB/s")
