"""Summarise an .ncu-rep (ncu --set full) into a small CSV that can be committed under profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_ncu_summary.csv"""
import csv, subprocess, sys, io, re
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["ID", "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic"]
idx = [hdr.index(w) for w in want if w in hdr]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
    for r in rows[2:]:
        r = list(r)
        ki = hdr.index("Kernel Name")
        r[ki] = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        w.writerow([r[i] for i in idx])
print("wrote", out, len(rows) - 2, "kernels")
