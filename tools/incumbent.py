#!/usr/bin/env python
"""Incumbent harness: times the kernel set the reference ACTUALLY runs on a GPU next to ours, same box, same inputs.

The reference has no native code: on a GPU its hot path resolves to third-party kernels -
  R-GPU-flex   HF Llama + make_flex_block_causal_mask + torch flex_attention (Inductor/Triton) + cuBLAS, eager
               RMSNorm / RoPE / SwiGLU (touchnet/bin/train.py:129-131 forces compile off with flex)
  R-GPU-liger  the same with Liger's Triton RMSNorm / SwiGLU (touchnet/models/llama/__init__.py:11-15)
This script measures them per op (and per decoder layer) at the bench shapes and prints ONE JSON object; bench.py embeds
it as `extras.incumbent` (never as the `--impl reference` arm, which stays the CPU path).  Every section is
failure-tolerant: what cannot run in the image is reported as {"error": "..."}.

    python tools/incumbent.py [--sections attn,gemm,norm,layer] [--out profiles/r02_incumbent.json]
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SCALE = 1 / math.sqrt(128)


def _time(fn, iters=10, warm=3, flush=None):
    """Median CUDA-event time of fn() in ms (L2 flushed between iterations when `flush` is a >L2 buffer)."""
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def _attn_flops(doc):
    """mask-exact forward FLOPs per head-dim-128 head: 4*hd*sum n_i(n_i+1)/2."""
    tot = 0
    for b in range(doc.shape[0]):
        ids = doc[b][doc[b] > 0]
        if ids.numel():
            n = torch.bincount(ids)[1:].double()
            tot += float((n * (n + 1) / 2).sum())
    return 4.0 * 128 * tot


def section_attention(dev, flush):
    from touchnet_b200 import batching, ops
    out = {}
    cases = {"cfg2_T8192_asr_docs": (1, 8192, 32, 8, "asr"), "cfg4_T32768_one_doc": (1, 32768, 32, 8, "one"),
             "cfg3_T4096_B2_mha": (2, 4096, 32, 32, "asr")}
    try:
        from torch.nn.attention.flex_attention import flex_attention
        from transformers.integrations.flex_attention import make_flex_block_causal_mask
        flex = torch.compile(flex_attention, dynamic=False)
    except Exception as e:
        flex, make_flex_block_causal_mask = None, None
        out["flex_error"] = f"{type(e).__name__}: {str(e)[:200]}"
    for name, (B, T, H, KV, kind) in cases.items():
        r = {}
        try:
            if kind == "asr":
                doc = batching.plan_audio_text_batch(2025, B, T, 128256, stride=4, max_s=30.0)[0]["attention_mask"].to(dev)
            else:
                doc = torch.ones(B, T, dtype=torch.int64, device=dev)
            g = torch.Generator(device="cpu").manual_seed(1)
            mk = lambda c: torch.randn(B * T, c * 128, generator=g).to(dev).bfloat16()
            q, k, v, do = mk(H), mk(KV), mk(KV), mk(H)
            fl = _attn_flops(doc) * H
            plan = ops.AttnPlan(doc)
            o, lse = ops.attn_fwd(q, k, v, plan, H, KV, SCALE)
            t_f = _time(lambda: ops.attn_fwd(q, k, v, plan, H, KV, SCALE), flush=flush)
            t_b = _time(lambda: ops.attn_bwd(q, k, v, o, do, lse, plan, H, KV, SCALE), flush=flush)
            r["ours_fwd_ms"], r["ours_bwd_ms"] = t_f, t_b
            r["ours_fwd_tflops_mask_exact"] = fl / t_f / 1e9
            r["ours_bwd_tflops_mask_exact_2p5x"] = 2.5 * fl / t_b / 1e9
            if flex is not None:
                bm = make_flex_block_causal_mask(doc)
                q4 = q.view(B, T, H, 128).transpose(1, 2).detach().requires_grad_(True)
                k4 = k.view(B, T, KV, 128).transpose(1, 2).detach().requires_grad_(True)
                v4 = v.view(B, T, KV, 128).transpose(1, 2).detach().requires_grad_(True)
                do4 = do.view(B, T, H, 128).transpose(1, 2)
                fwd = lambda: flex(q4, k4, v4, block_mask=bm, enable_gqa=(H != KV), scale=SCALE)
                of = fwd()
                t_ff = _time(lambda: fwd(), flush=flush)

                def fb():
                    q4.grad = k4.grad = v4.grad = None
                    fwd().backward(do4)
                t_fb = _time(fb, flush=flush)
                r["flex_compiled_fwd_ms"], r["flex_compiled_bwd_ms"] = t_ff, max(t_fb - t_ff, 1e-6)
                r["flex_fwd_tflops_mask_exact"] = fl / t_ff / 1e9
                r["ours_over_flex_fwd"] = t_ff / t_f
                r["ours_over_flex_bwd"] = r["flex_compiled_bwd_ms"] / t_b
                del of
            r["mask_exact_fwd_tflop"] = fl / 1e12
        except Exception as e:
            r["error"] = f"{type(e).__name__}: {str(e)[:300]}"
        out[name] = r
    return out


def section_gemm(dev, flush):
    """cuBLAS (torch.matmul, bf16 in, fp32 accumulate) vs tn_gemm_bf16 on the Llama-3-8B shapes, M = 8192 tokens."""
    from touchnet_b200 import ops
    out = {}
    shapes = {"qkv": (8192, 6144, 4096), "o_proj": (8192, 4096, 4096), "gate_up": (8192, 14336, 4096),
              "down": (8192, 4096, 14336), "lm_head": (8192, 128256, 4096)}
    for name, (M, N, K) in shapes.items():
        r = {}
        try:
            a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
            b = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
            dy = (torch.randn(M, N, device=dev) * 0.1).bfloat16()
            fl = 2.0 * M * N * K
            for tag, ours, theirs in (
                    ("fwd", lambda: ops.gemm(a, b), lambda: torch.matmul(a, b.t())),
                    ("dgrad", lambda: ops.gemm(dy, b, b_mn=True), lambda: torch.matmul(dy, b)),
                    ("wgrad_bf16", lambda: ops.gemm(dy, a, a_mn=True, b_mn=True), lambda: torch.matmul(dy.t(), a))):
                t_o, t_c = _time(ours, flush=flush), _time(theirs, flush=flush)
                r[tag] = {"ours_ms": t_o, "cublas_ms": t_c, "ours_tflops": fl / t_o / 1e9, "cublas_tflops": fl / t_c / 1e9,
                          "ours_over_cublas": t_c / t_o}
        except Exception as e:
            r["error"] = f"{type(e).__name__}: {str(e)[:300]}"
        out[name] = r
        torch.cuda.empty_cache()
    return out


def section_norm(dev, flush):
    """RMSNorm fwd+bwd and SwiGLU (gate/up outputs given) at [8192, 4096] / [8192, 14336]: HF eager, Liger Triton, ours."""
    from touchnet_b200 import ops
    out = {}
    rows, d, ffn = 8192, 4096, 14336
    x = torch.randn(rows, d, device=dev).bfloat16()
    w = torch.ones(d, device=dev)
    dy = torch.randn(rows, d, device=dev).bfloat16()
    r = {}
    try:
        y, _, rstd = ops.rmsnorm_fwd(x, w, 1e-5)
        r["ours_fwd_ms"] = _time(lambda: ops.rmsnorm_fwd(x, w, 1e-5), flush=flush)
        r["ours_bwd_ms"] = _time(lambda: ops.rmsnorm_bwd(x, dy, w, rstd), flush=flush)
        r["ours_fwd_gbps"] = rows * d * 4 / r["ours_fwd_ms"] / 1e6
        r["ours_bwd_gbps"] = rows * d * 6 / r["ours_bwd_ms"] / 1e6
    except Exception as e:
        r["ours_error"] = f"{type(e).__name__}: {str(e)[:200]}"
    try:
        from transformers.models.llama.modeling_llama import LlamaRMSNorm
        hf = LlamaRMSNorm(d, 1e-5).to(dev).bfloat16()
        xr = x.clone().requires_grad_(True)
        r["hf_eager_fwd_ms"] = _time(lambda: hf(xr), flush=flush)

        def fb():
            xr.grad = None
            hf(xr).backward(dy)
        r["hf_eager_fwd_bwd_ms"] = _time(fb, flush=flush)
    except Exception as e:
        r["hf_error"] = f"{type(e).__name__}: {str(e)[:200]}"
    try:
        from liger_kernel.transformers.rms_norm import LigerRMSNorm
        lg = LigerRMSNorm(d, eps=1e-5).to(dev).bfloat16()
        xr2 = x.clone().requires_grad_(True)
        r["liger_fwd_ms"] = _time(lambda: lg(xr2), flush=flush)

        def fb2():
            xr2.grad = None
            lg(xr2).backward(dy)
        r["liger_fwd_bwd_ms"] = _time(fb2, flush=flush)
    except Exception as e:
        r["liger_error"] = f"{type(e).__name__}: {str(e)[:200]}"
    out["rmsnorm_8192x4096"] = r
    s = {}
    try:
        g = torch.randn(rows, ffn, device=dev).bfloat16()
        u = torch.randn(rows, ffn, device=dev).bfloat16()
        dh = torch.randn(rows, ffn, device=dev).bfloat16()
        s["ours_swiglu_bwd_ms"] = _time(lambda: ops.swiglu_bwd(g, u, dh), flush=flush)
        gr, ur = g.clone().requires_grad_(True), u.clone().requires_grad_(True)

        def eager():
            gr.grad = ur.grad = None
            (torch.nn.functional.silu(gr) * ur).backward(dh)
        s["eager_silu_mul_fwd_bwd_ms"] = _time(eager, flush=flush)
        try:
            from liger_kernel.ops.swiglu import LigerSiLUMulFunction

            def lig():
                gr.grad = ur.grad = None
                LigerSiLUMulFunction.apply(gr, ur).backward(dh)
            s["liger_silu_mul_fwd_bwd_ms"] = _time(lig, flush=flush)
        except Exception as e:
            s["liger_error"] = f"{type(e).__name__}: {str(e)[:200]}"
    except Exception as e:
        s["error"] = f"{type(e).__name__}: {str(e)[:200]}"
    out["swiglu_8192x14336"] = s
    return out


def section_layer(dev, flush):
    """One Llama-3-8B decoder layer forward+backward on a packed T=8192 row: HF (bf16 params, flex_attention BlockMask,
    cuBLAS, eager norms/rope/SwiGLU = R-GPU-flex) vs the B200 block, both as 2-layer models with a tiny vocabulary so the
    embedding / lm_head are negligible; ms per layer = total / 2."""
    from types import SimpleNamespace as NS
    from touchnet_b200 import batching, modeling
    out = {}
    B, T, L = 1, 8192, 2
    rs = {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
          "original_max_position_embeddings": 8192}
    buf = batching.plan_audio_text_batch(2025, B, T, 128256, stride=4, max_s=30.0)[0]
    doc, pos = buf["attention_mask"].to(dev), buf["position_ids"].to(dev)
    ids = torch.randint(1, 256, (B, T), device=dev)
    try:
        cfg = NS(hidden_size=4096, intermediate_size=14336, num_hidden_layers=L, num_attention_heads=32,
                 num_key_value_heads=8, head_dim=128, vocab_size=256, rms_norm_eps=1e-5, rope_theta=500000.0,
                 rope_scaling=rs, attention_bias=False, tie_word_embeddings=False, initializer_range=0.02,
                 model_type="llama", pad_token_id=0)
        ours = modeling.B200LlamaForCausalLM(cfg).to(dev)
        ours.post_init()
        ours.to(torch.bfloat16)           # bf16 parameters, as FSDP2's mixed-precision all-gather hands them to the blocks

        def step_ours():
            ours.zero_grad(set_to_none=True)
            ours(input_ids=ids, attention_mask=doc, position_ids=pos).logits.float().mean().backward()
        out["ours_ms_per_layer_fwd_bwd"] = _time(step_ours, iters=5, flush=flush) / L
        del ours
    except Exception as e:
        out["ours_error"] = f"{type(e).__name__}: {str(e)[:300]}"
    torch.cuda.empty_cache()
    try:
        from transformers import LlamaConfig, LlamaForCausalLM
        from transformers.integrations.flex_attention import make_flex_block_causal_mask
        hc = LlamaConfig(hidden_size=4096, intermediate_size=14336, num_hidden_layers=L, num_attention_heads=32,
                         num_key_value_heads=8, head_dim=128, vocab_size=256, rms_norm_eps=1e-5, rope_theta=500000.0,
                         rope_scaling=dict(rs), tie_word_embeddings=False, attention_bias=False)
        hc._attn_implementation = "flex_attention"
        hf = LlamaForCausalLM(hc).to(dev).to(torch.bfloat16).train()
        bm = make_flex_block_causal_mask(doc)

        def step_hf():
            hf.zero_grad(set_to_none=True)
            hf(input_ids=ids, attention_mask=bm, position_ids=pos).logits.float().mean().backward()
        out["hf_flex_cublas_eager_ms_per_layer_fwd_bwd"] = _time(step_hf, iters=5, flush=flush) / L
        if "ours_ms_per_layer_fwd_bwd" in out:
            out["ours_over_hf_flex"] = out["hf_flex_cublas_eager_ms_per_layer_fwd_bwd"] / out["ours_ms_per_layer_fwd_bwd"]
        try:
            from liger_kernel.transformers import apply_liger_kernel_to_llama
            apply_liger_kernel_to_llama(rope=True, rms_norm=True, swiglu=True, cross_entropy=False,
                                        fused_linear_cross_entropy=False, model=hf)
            out["hf_flex_liger_ms_per_layer_fwd_bwd"] = _time(step_hf, iters=5, flush=flush) / L
            if "ours_ms_per_layer_fwd_bwd" in out:
                out["ours_over_hf_liger"] = out["hf_flex_liger_ms_per_layer_fwd_bwd"] / out["ours_ms_per_layer_fwd_bwd"]
        except Exception as e:
            out["liger_error"] = f"{type(e).__name__}: {str(e)[:300]}"
        del hf
    except Exception as e:
        out["hf_error"] = f"{type(e).__name__}: {str(e)[:300]}"
    torch.cuda.empty_cache()
    return out


SECTIONS = {"attn": section_attention, "gemm": section_gemm, "norm": section_norm, "layer": section_layer}


def measure(sections=("attn", "gemm", "norm", "layer"), budget_s: float = 600.0) -> dict:
    dev = torch.device("cuda", torch.cuda.current_device())
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2
    res = {"note": "incumbent = the third-party kernels the reference's GPU path resolves to (SURVEY 2.2 K1-K7), timed "
                   "with CUDA events, median of 10 after 3 warm-ups, L2 flushed between iterations; mask-exact FLOPs"}
    t0 = time.time()
    for s in sections:
        if time.time() - t0 > budget_s:
            res[s] = {"skipped": "time budget"}
            continue
        try:
            res[s] = SECTIONS[s](dev, flush)
        except Exception as e:       # a section must never take the caller down
            res[s] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
        torch.cuda.empty_cache()
    res["seconds"] = round(time.time() - t0, 1)
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--sections", default="attn,gemm,norm,layer")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    r = measure(tuple(a.sections.split(",")))
    txt = json.dumps(r, indent=1)
    print(txt)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        open(a.out, "w").write(txt + "\n")
