"""Development check (GPU): tcgen05 GEMM variants vs torch.matmul.  Each case runs in its own subprocess with a
timeout so a protocol bug cannot hang the box.  Usage: python tools/check_gemm.py [--case NAME]"""
import argparse, json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = {
    # name: (M, N, K, a_mn, b_mn, d_f32, residual, swiglu)
    "nt_1tile":   (128, 256, 64, 0, 0, 0, 0, 0),
    "nt_k4":      (128, 256, 256, 0, 0, 0, 0, 0),
    "nt_multi":   (512, 1024, 512, 0, 0, 0, 0, 0),
    "nt_narrow":  (256, 128, 256, 0, 0, 0, 0, 0),
    "nn_1tile":   (128, 256, 64, 0, 1, 0, 0, 0),
    "nn_multi":   (512, 1024, 512, 0, 1, 0, 0, 0),
    "tn_1tile":   (128, 256, 64, 1, 1, 0, 0, 0),
    "tn_multi":   (512, 1024, 512, 1, 1, 0, 0, 0),
    "tn_f32":     (512, 1024, 512, 1, 1, 1, 0, 0),
    "tn_f32_acc": (512, 1024, 512, 1, 1, 1, 1, 0),
    "nt_res":     (512, 1024, 512, 0, 0, 0, 1, 0),
    "nt_odd":     (200, 264, 1040, 0, 0, 0, 0, 0),
    "nn_odd":     (200, 264, 1040, 0, 1, 0, 0, 0),
    "tn_odd":     (264, 1040, 200, 1, 1, 1, 0, 0),
    "swiglu":     (512, 1024, 512, 0, 0, 0, 0, 1),
    "swiglu_odd": (200, 264, 1040, 0, 0, 0, 0, 1),
    "pair_tail":  (300, 520, 1040, 0, 0, 0, 1, 0),
    "pair_tail_nn": (300, 520, 1040, 0, 1, 0, 0, 0),
    "pair_tail_tn": (520, 1040, 300, 1, 1, 1, 1, 0),
    "swiglu_tail": (300, 392, 1040, 0, 0, 0, 0, 1),
    "perf_qo":    (8192, 4096, 4096, 0, 0, 0, 0, 0),
    "perf_kv":    (8192, 1024, 4096, 0, 0, 0, 0, 0),
    "perf_lmhead": (8192, 128256, 4096, 0, 0, 0, 0, 0),
    "perf_wgrad_qo": (4096, 4096, 8192, 1, 1, 1, 0, 0),
    "perf_down":  (8192, 4096, 14336, 0, 0, 0, 1, 0),
    "perf_dgrad": (8192, 4096, 14336, 0, 1, 0, 0, 0),
    "perf_wgrad": (14336, 4096, 8192, 1, 1, 1, 0, 0),
    "perf_swiglu": (8192, 14336, 4096, 0, 0, 0, 0, 1),
}


def run_case(name):
    import torch
    from touchnet_b200 import _lib
    M, N, K, a_mn, b_mn, d_f32, res, swiglu = CASES[name]
    torch.manual_seed(0)
    dev = "cuda"
    st = _lib.stream_handle()
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    out = {"case": name, "shape": [M, N, K]}
    if swiglu:
        Wg = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        Wu = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        G = torch.empty(M, N, device=dev, dtype=torch.bfloat16); U = torch.empty_like(G); H = torch.empty_like(G)
        def f():
            _lib.call("tn_gemm_swiglu_bf16", A.data_ptr(), K, Wg.data_ptr(), Wu.data_ptr(), K, G.data_ptr(), U.data_ptr(), H.data_ptr(), N, M, N, K, st)
        f(); torch.cuda.synchronize()
        g_ref = (A.float() @ Wg.float().t()); u_ref = (A.float() @ Wu.float().t())
        gb, ub = g_ref.bfloat16(), u_ref.bfloat16()
        h_ref = (torch.nn.functional.silu(gb.float()).bfloat16().float() * ub.float())
        out["err_g"] = (G.float() - g_ref).abs().max().item(); out["err_u"] = (U.float() - u_ref).abs().max().item()
        out["err_h"] = (H.float() - h_ref).abs().max().item(); out["ref_max"] = h_ref.abs().max().item()
        ok = out["err_g"] < 0.05 * g_ref.abs().max().item() + 1e-2 and out["err_h"] < 0.05 * h_ref.abs().max().item() + 1e-2
        flops = 2.0 * M * (2 * N) * K
    else:
        B = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        A_store = A.t().contiguous() if a_mn else A          # [K,M] if a_mn
        B_store = B.t().contiguous() if b_mn else B          # [K,N] if b_mn
        lda = M if a_mn else K
        ldb = N if b_mn else K
        D = torch.zeros(M, N, device=dev, dtype=torch.float32 if d_f32 else torch.bfloat16)
        R = None
        if res:
            R = (torch.randn(M, N, device=dev)).to(D.dtype)
        def f():
            _lib.call("tn_gemm_bf16", A_store.data_ptr(), lda, a_mn, B_store.data_ptr(), ldb, b_mn, D.data_ptr(), N, d_f32,
                      R.data_ptr() if res else None, N, M, N, K, st)
        f(); torch.cuda.synchronize()
        ref = A.float() @ B.float().t()
        if res:
            ref = (ref.bfloat16().float() if not d_f32 else ref) + R.float()
        err = (D.float() - ref).abs().max().item()
        out["err"] = err; out["ref_max"] = ref.abs().max().item()
        ok = err < 0.02 * ref.abs().max().item() + 1e-2
        if not ok:
            bad = ((D.float() - ref).abs() > 0.02 * ref.abs().max().item() + 1e-2)
            idx = bad.nonzero()
            out["n_bad"] = int(bad.sum().item()); out["first_bad"] = idx[:4].tolist()
            out["bad_rows"] = sorted(set((idx[:, 0] // 8 * 8).tolist()))[:16]
            out["bad_cols"] = sorted(set((idx[:, 1] // 8 * 8).tolist()))[:16]
        flops = 2.0 * M * N * K
    out["ok"] = bool(ok)
    if name.startswith("perf"):
        for _ in range(3): f()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        iters = 20
        for _ in range(iters): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        out["ms"] = ms; out["tflops"] = flops / ms / 1e9
        if not swiglu:
            # cuBLAS for comparison
            Bm = B
            for _ in range(3): torch.matmul(A, Bm.t())
            torch.cuda.synchronize(); e0.record()
            for _ in range(iters): torch.matmul(A, Bm.t())
            e1.record(); torch.cuda.synchronize()
            out["cublas_nt_tflops"] = flops / (e0.elapsed_time(e1) / iters) / 1e9
    print("RESULT " + json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--case", default=None); ap.add_argument("--only", default=None)
    ap.add_argument("--inproc", action="store_true", help="all cases in this process (one torch import)")
    a = ap.parse_args()
    if a.case:
        run_case(a.case); return
    if a.inproc:
        import traceback
        for n in CASES:
            if a.only is None or a.only in n:
                try:
                    run_case(n)
                except Exception as e:
                    print("RESULT " + json.dumps({"case": n, "ok": False, "exc": repr(e)[:300]}), flush=True)
        print("DONE", flush=True); return
    os.environ.setdefault("TN_DEV_PARTIAL", "1")
    names = [n for n in CASES if (a.only is None or a.only in n)]
    nfail = 0
    for n in names:
        t = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--case", n], capture_output=True, text=True, timeout=120)
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if lines: print(lines[-1][7:])
            else:
                nfail += 1
                print(json.dumps({"case": n, "ok": False, "rc": r.returncode, "stdout": r.stdout[-1500:], "stderr": r.stderr[-1500:]}))
        except subprocess.TimeoutExpired:
            nfail += 1
            print(json.dumps({"case": n, "ok": False, "timeout": True}))
        sys.stdout.flush()
    print("DONE", flush=True)

if __name__ == "__main__":
    main()
