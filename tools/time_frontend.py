"""Time the GPU fbank kernel on a bench-sized batch (CUDA events) and report achieved algorithmic GB/s."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from touchnet_b200 import frontend
g = torch.Generator().manual_seed(0)
lens = [int(16000 * (1 + 29 * float(torch.rand((), generator=g)))) for _ in range(64)]      # ~16 min of audio
wav = (0.3 * (2 * torch.rand(sum(lens), generator=g) - 1)).cuda()
for _ in range(3): out, frames = frontend.fbank_batch(wav, lens)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(10):
    flush.zero_()                       # flush L2 between timed iterations
    e0.record(); out, frames = frontend.fbank_batch(wav, lens); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
alg = wav.numel() * 4 + out.numel() * 4
print(f"fbank {'DFT' if os.environ.get('TN_FBANK_DFT') else 'FFT'}: {sum(frames)} frames, {ms*1e3:.1f} us, algorithmic {alg/1e6:.1f} MB -> {alg/ms/1e6:.1f} GB/s, {sum(frames)/ms/1e3:.1f} Mframes/s")
