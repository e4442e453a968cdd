"""Text summary of one .ncu-rep capture (raw metrics + hottest CUDA source lines) for profiles/.

    python tools/ncu_kernel_summary.py <capture.ncu-rep> <out.txt> "<header line>"
"""
import csv
import subprocess
import sys

rep, out, header = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
d = {h: (u, v) for h, u, v in zip(rr[0], rr[1], rr[2])}
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max"]
with open(out, "w") as f:
    f.write("# %s\n" % header)
    for k in keys:
        if k in d:
            f.write("%-92s %-16s %s\n" % (k, d[k][0], d[k][1]))
    for h in sorted(d):
        if "issue_stalled" in h and "per_issue_active" in h:
            try:
                if float(d[h][1]) > 0.15:
                    f.write("%-92s %-16s %s\n" % (h, d[h][0], d[h][1]))
            except ValueError:
                pass
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    tmp = out + ".src.csv"
    open(tmp, "w").write(src)
    lines = subprocess.run([sys.executable, __file__.replace("ncu_kernel_summary.py", "ncu_lines.py"), tmp, "24"], capture_output=True, text=True).stdout
    f.write("\n# hottest CUDA source lines (warp-state samples, executed warp instructions)\n" + lines)
import os
os.remove(out + ".src.csv")
