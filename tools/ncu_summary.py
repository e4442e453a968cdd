"""Turns the ncu artefacts of tools/gpu_profile.sh into the text summaries committed under profiles/.

    python tools/ncu_summary.py gpurun_out profiles r01 <env_steps_in_profiled_launch>
"""
import csv
import json
import subprocess
import sys
from collections import defaultdict

src, dst, tag = sys.argv[1], sys.argv[2], sys.argv[3]
steps = float(sys.argv[4]) if len(sys.argv) > 4 else None

# ---- launch list ----------------------------------------------------------------------------------------
rows = list(csv.reader(l for l in open("%s/launches.csv" % src) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e6 if r[ui] in ("ns", "nsecond") else (v / 1e3 if r[ui] in ("us", "usecond") else v)
    agg[r[ki]][0] += 1
    agg[r[ki]][1] += v
tot = sum(a[1] for a in agg.values())
with open("%s/%s_launches.txt" % (dst, tag), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 1 --warmup 1 --no-cpu-baseline\n")
    f.write("# per-kernel launch count, summed device time, share of all launches (cold-cache, serialised: compare shares)\n")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write("%-110s n=%3d  %10.3f ms  %5.1f%%\n" % (k[:110], a[0], a[1], 100 * a[1] / tot))

# ---- full capture ---------------------------------------------------------------------------------------
raw = subprocess.run(["ncu", "-i", "%s/prof_inner.ncu-rep" % src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
d = {h: (u, v) for h, u, v in zip(rr[0], rr[1], rr[2])}
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max"]


def num(k):
    u, v = d[k]
    x = float(v.replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u)
    return x * scale if scale else x


with open("%s/%s_inner_loop_kernel.txt" % (dst, tag), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:inner_loop_kernel -c 1  python bench.py --steps 1 --warmup 1\n")
    f.write("# kernel: %s\n" % rr[2][rr[0].index("Kernel Name")] if "Kernel Name" in rr[0] else "")
    for k in keys:
        if k in d:
            f.write("%-90s %-16s %s\n" % (k, d[k][0], d[k][1]))
    for h in sorted(d):
        if "issue_stalled" in h and "per_issue_active" in h and float(d[h][1] or 0) > 0.04:
            f.write("%-90s %-16s %s\n" % (h, d[h][0], d[h][1]))
    if steps:
        traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
        f.write("\n# env steps in this launch: %d\n" % steps)
        f.write("# DRAM traffic per launch: %.3f GB  = %.0f B per env step (algorithmic replay bytes: 8800 B/step)\n" % (traffic / 1e9, traffic / steps))
        f.write("# warp instructions per env step: %.0f\n" % (num("smsp__inst_executed.sum") / steps))
        json.dump({"kernel": "inner_loop_kernel<4,2,2,tanh>", "dram_bytes_per_launch": traffic, "env_steps_per_launch": steps,
                   "dram_bytes_per_env_step": traffic / steps, "duration_ms": num("gpu__time_duration.sum") if d["gpu__time_duration.sum"][0] == "ms" else None,
                   "source": "%s/%s_inner_loop_kernel.txt" % (dst, tag)}, open("%s/%s_traffic.json" % (dst, tag), "w"), indent=1)

# ---- opcode mix -----------------------------------------------------------------------------------------
srcp = subprocess.run(["ncu", "-i", "%s/prof_inner.ncu-rep" % src, "--page", "source", "--csv"], capture_output=True, text=True).stdout
open("/tmp/_src.csv", "w").write(srcp)
mix = subprocess.run([sys.executable, "tools/ncu_opmix.py", "/tmp/_src.csv", str(steps or 1)], capture_output=True, text=True).stdout
open("%s/%s_inner_loop_opmix.txt" % (dst, tag), "w").write("# SASS opcode mix of the profiled launch (ncu --page source), executed warp instructions\n" + mix)
print("wrote summaries to", dst)
