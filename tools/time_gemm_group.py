"""Raster-group sweep of the CTA-pair GEMM on the Llama-3-8B shapes (tn_set_gemm_group), operands rotated over more
buffers than fit in L2.  Prints TFLOP/s per (shape, group)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from touchnet_b200 import _lib, ops
dev = torch.device("cuda")
M = 8192
SHAPES = {  # name: (a_shape, b_shape, a_mn, b_mn, out_f32)   D = op(A) op(B)
    "down fwd   [8192x4096x14336]": ((M, 14336), (4096, 14336), False, False, False),
    "gate dgrad [8192x4096x14336]": ((M, 14336), (14336, 4096), False, True, False),
    "gate wgrad [14336x4096x8192]": ((M, 14336), (M, 4096), True, True, True),
    "qkv-ish fwd [8192x4096x4096]": ((M, 4096), (4096, 4096), False, False, False),
    "lm_head fwd [8192x128256x4096]": ((M, 4096), (128256, 4096), False, False, False),
}
NB = 3
def flops(a, b, a_mn, b_mn):
    m, k = (a[1], a[0]) if a_mn else a
    n = b[1] if b_mn else b[0]
    return 2.0 * m * n * k
for name, (ash, bsh, a_mn, b_mn, f32) in SHAPES.items():
    As = [torch.randn(ash, device=dev).bfloat16() for _ in range(NB)]
    Bs = [(torch.randn(bsh, device=dev) * 0.05).bfloat16() for _ in range(NB if bsh[0] < 100000 else 1)]
    line = f"{name:32s}"
    for grp in (2, 4, 8, 16, 32):
        _lib.call("tn_set_gemm_group", grp)
        f = lambda i: ops.gemm(As[i % NB], Bs[i % len(Bs)], a_mn=a_mn, b_mn=b_mn, out_f32=f32)
        for i in range(3): f(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        it = 12
        e0.record()
        for i in range(it): f(i)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / it
        line += f"  g{grp}: {flops(ash, bsh, a_mn, b_mn) / ms / 1e9:6.0f}"
    print(line, "TF/s")
_lib.call("tn_set_gemm_group", 8)
