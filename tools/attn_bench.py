#!/usr/bin/env python
"""Stand-alone timing of the packed-attention kernels at the BASELINE shapes (used directly and under ncu).
    python tools/attn_bench.py [--case cfg2|cfg3|cfg4] [--iters N]"""
import argparse, json, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from touchnet_b200 import batching, ops
from tools.incumbent import _attn_flops, _time

CASES = {"cfg2": (1, 8192, 32, 8, "asr"), "cfg3": (2, 4096, 32, 32, "asr"), "cfg4": (1, 32768, 32, 8, "one"),
         "cfg5": (1, 16384, 32, 8, "asr")}
ap = argparse.ArgumentParser()
ap.add_argument("--case", default="cfg2,cfg4")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--hd", type=int, default=128)
a = ap.parse_args()
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peak = 1449.3
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
except Exception:
    pass
for name in a.case.split(","):
    B, T, H, KV, kind = CASES[name]
    doc = (batching.plan_audio_text_batch(2025, B, T, 128256, stride=4, max_s=30.0)[0]["attention_mask"] if kind == "asr"
           else torch.ones(B, T, dtype=torch.int64)).to(dev)
    g = torch.Generator(device="cpu").manual_seed(1)
    mk = lambda c: torch.randn(B * T, c * 128, generator=g).to(dev).bfloat16()
    q, k, v, do = mk(H), mk(KV), mk(KV), mk(H)
    sc = 1 / math.sqrt(128)
    plan = ops.AttnPlan(doc)
    o, lse = ops.attn_fwd(q, k, v, plan, H, KV, sc)
    fl = _attn_flops(doc) * H
    nblk = (T + 127) // 128
    meta = plan.meta[: B * nblk * 4].view(B * nblk, 4).cpu()
    tiles = int((meta[:, 1] - meta[:, 0]).clamp(min=0).sum()) * H
    tile_fl = tiles * 4.0 * 128 * 128 * 128
    tf = _time(lambda: ops.attn_fwd(q, k, v, plan, H, KV, sc), iters=a.iters, flush=flush)
    tb = _time(lambda: ops.attn_bwd(q, k, v, o, do, lse, plan, H, KV, sc), iters=a.iters, flush=flush)
    print(json.dumps({"case": name, "fwd_ms": tf, "bwd_ms": tb, "fwd_tflops_mask_exact": fl / tf / 1e9,
                      "bwd_tflops_mask_exact_2p5x": 2.5 * fl / tb / 1e9, "fwd_mfu_mask_exact": fl / tf / 1e9 / peak,
                      "bwd_mfu_mask_exact": 2.5 * fl / tb / 1e9 / peak, "fwd_mfu_tile_granular": tile_fl / tf / 1e9 / peak,
                      "bwd_mfu_tile_granular": 2.5 * tile_fl / tb / 1e9 / peak, "tiles": tiles,
                      "mask_exact_over_tile": fl / tile_fl, "peak_tflops_sustained": peak}))
