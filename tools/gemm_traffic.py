"""profiles/rNN_gemm_traffic.json from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
-k regex:gemm --csv` capture of `bench.py --layers 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline` (tools/gpu_round.sh).

DRAM bytes are measured per GEMM launch of the LAST step in the capture; the 2 profiled decoder blocks are scaled to the
32 of the named workload (the 5 launches outside the blocks - projector fwd / wgrad, lm_head fwd / dgrad / wgrad - are
counted once).  Algorithmic bytes per launch = every operand read once + the output written once, from the GEMM inventory
of one step (M = B*T = 8192 rows).  usage: gemm_traffic.py capture.csv out.json"""
import csv, json, sys

M, D, NQ, NKV, FFN, V, F = 8192, 4096, 4096, 1024, 14336, 128256, 400
BF, F32 = 2, 4


def g(m, n, k, out=BF, extra=0):
    """algorithmic bytes of D[m,n] = A[m,k] B[n,k]^T: read A, read B, write D (+ extra operands, e.g. a residual)."""
    return (m * k + n * k) * BF + m * n * out + extra


QKV = NQ + 2 * NKV
LAYER_FWD = [g(M, QKV, D, extra=2 * M * 64 * BF),          # fused QKV projection + RoPE tables
             g(M, D, NQ, extra=M * D * BF),                 # o_proj + residual
             g(M, 2 * FFN, D) + M * FFN * BF,               # gate/up with SwiGLU epilogue: writes G, U and H
             g(M, D, FFN, extra=M * D * BF)]                # down_proj + residual
LAYER_BWD = [g(M, FFN, D),                                  # down dgrad
             g(D, FFN, M, out=F32),                         # down wgrad (fp32 master gradient)
             g(M, D, FFN), g(M, D, FFN, extra=M * D * BF),  # gate dgrad, up dgrad (+accumulate)
             g(FFN, D, M, out=F32), g(FFN, D, M, out=F32),  # gate / up wgrad
             g(M, NQ, D), g(D, NQ, M, out=F32),             # o dgrad, wgrad
             g(M, D, QKV), g(QKV, D, M, out=F32)]           # fused QKV dgrad, wgrad
NCH, MC = 2, 4096                                              # fused lm_head + loss: chunks of 4096 token rows (loss.FUSED_CE_CHUNK)
OUTSIDE = [g(M, D, F), g(D, F, M, out=F32)] + [
    x for c in range(NCH) for x in (g(MC, V, D), g(MC, D, V), g(V, D, MC, out=F32, extra=(V * D * F32 if c else 0)))]
L = 32
ALG_TOTAL = L * (sum(LAYER_FWD) + sum(LAYER_BWD)) + sum(OUTSIDE)
N_LAUNCH = L * (len(LAYER_FWD) + len(LAYER_BWD)) + len(OUTSIDE)


def main(path, out):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    per = {}
    for r in rows[hi + 1:]:
        if len(r) <= vi or "gemm" not in r[ki]:
            continue
        d = per.setdefault(int(r[ii]), {})
        d[r[mi]] = float(r[vi].replace(",", ""))
    launches = [per[k] for k in sorted(per)]
    per_step = 2 * 14 + 2 + 3 * NCH
    assert len(launches) >= per_step and len(launches) % per_step == 0, (len(launches), per_step)
    last = launches[-per_step:]
    dram = [x["dram__bytes_read.sum"] + x["dram__bytes_write.sum"] for x in last]
    # launch order inside a step: projector fwd | 2 x 4 block fwd | NCH x (lm_head fwd, dgrad, wgrad) | 2 x 10 block bwd |
    # projector wgrad
    h0 = 9 + 3 * NCH
    block = sum(dram[1:9]) + sum(dram[h0:h0 + 20])
    outside = dram[0] + sum(dram[9:h0]) + dram[h0 + 20]
    total = block * (L / 2) + outside
    res = {"source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:gemm on bench.py --layers 2 ({path}), "
                     "block launches scaled to 32 layers (tools/gemm_traffic.py)",
           "gemm_launches_per_step": N_LAUNCH,
           "avg_dram_bytes_per_gemm_launch": total / N_LAUNCH,
           "avg_algorithmic_bytes_per_gemm_launch": ALG_TOTAL / N_LAUNCH,
           "dram_over_algorithmic": total / ALG_TOTAL,
           "total_dram_GB_per_step": total / 1e9, "total_algorithmic_GB_per_step": ALG_TOTAL / 1e9}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
