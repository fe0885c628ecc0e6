#!/usr/bin/env python
"""Turns the files tools/gpu_final.sh brings back in gpurun_out/ into the tracked summaries:
usage: profiles/summarize.py <gpurun tag, e.g. r01c> <round prefix, e.g. r01>
Needs the librtx.so the capture was taken with (nvdisasm line info)."""
import csv
import json
import os
import subprocess
import sys

tag, rnd = sys.argv[1:3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]

rows = list(csv.reader(open(os.path.join(G, tag + "_k_render_raw.csv"))))
hdr, units, vals = rows[0], rows[1], rows[2]
out = ["# ncu --set full --clock-control none --import-source on, k_render<false>, one launch",
       "# command: python bench.py --steps 1 --warmup 1 --spp 32 --no-cpu  (1200x800, reference mesh mix)"]
for k in KEYS:
    if k in hdr:
        i = hdr.index(k)
        out.append("%-70s %-16s %s" % (k, units[i], vals[i]))
# the metrics-only pass at the bench configuration
t = {}
for r in csv.reader(open(os.path.join(G, tag + "_k_render_traffic.csv"))):
    if len(r) > 14 and r[12] in ("dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum", "gpu__time_duration.sum"):
        t[r[12]] = int(float(r[14]))
out.append("")
out.append("# metrics-only pass at the bench configuration (500 spp, one launch = one frame):")
for k, v in sorted(t.items()):
    out.append("%-70s %d" % (k, v))
open(os.path.join(P, rnd + "_k_render_ncu_summary.txt"), "w").write("\n".join(out) + "\n")
json.dump({"kernel": "k_render", "config": "1200x800, 500 spp, depth 50, mesh 9/6/3/8/6/3",
           "dram_bytes_read": t["dram__bytes_read.sum"], "dram_bytes_write": t["dram__bytes_write.sum"],
           "lts_t_bytes": t["lts__t_bytes.sum"], "l1tex_t_bytes": t["l1tex__t_bytes.sum"], "gpu_time_ns": t["gpu__time_duration.sum"],
           "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__t_bytes.sum,gpu__time_duration.sum --clock-control none (tools/gpu_final.sh)"},
          open(os.path.join(P, rnd + "_k_render_traffic.json"), "w"), indent=1)
# hot lines and stalls: join the source page with nvdisasm line info
sass = os.path.join(G, tag + "_librtx.sass")
lib = os.path.join(ROOT, "rtxplay_b200", "librtx.so")
cub = os.path.join(G, tag + "_librtx.cubin")
subprocess.check_call("cd %s && rm -f *.cubin && cuobjdump -xelf all %s >/dev/null && mv $(ls -S *.cubin | head -1) %s.keep && rm -f *.cubin && mv %s.keep %s" % (G, lib, cub, cub, cub), shell=True)
subprocess.check_call("nvdisasm -g -c %s > %s" % (cub, sass), shell=True)
src = os.path.join(G, tag + "_k_render_source.csv")
for script, name in (("sass_lines.py", "hot_lines"), ("ncu_stalls.py", "stalls")):
    txt = subprocess.check_output([sys.executable, os.path.join(P, script), src, sass, "k_renderILb0"]).decode()
    open(os.path.join(P, "%s_k_render_%s.txt" % (rnd, name)), "w").write(txt)
for a, b in (("_bench.json", "_bench.json"), ("_bench_reference.json", "_bench_reference.json"), ("_launches.csv", "_launches.csv")):
    open(os.path.join(P, rnd + b), "w").write(open(os.path.join(G, tag + a)).read())
print("\n".join(out))
