#!/usr/bin/env python
"""Turns the files tools/gpu_final.sh brings back in gpurun_out/ into the tracked round-2 summaries.
usage: profiles/summarize_r02.py <gpurun tag, e.g. r02>   (needs the librtx.so the captures were taken with)"""
import csv
import json
import os
import shutil
import subprocess
import sys

tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
lib = os.path.join(ROOT, "rtxplay_b200", "librtx.so")
cub = os.path.join(G, tag + "_librtx.cubin")
sass = os.path.join(G, tag + "_librtx.sass")
subprocess.check_call("cd %s && rm -f *.cubin && cuobjdump -xelf all %s >/dev/null && mv $(ls -S *.cubin | head -1) %s.keep && rm -f *.cubin && mv %s.keep %s" % (G, lib, cub, cub, cub), shell=True)
subprocess.check_call("nvdisasm -g -c %s > %s" % (cub, sass), shell=True)
for kern, name, mangled in (("reg", "k_render", "k_renderILb0"), ("q", "k_render_q", "k_render_qILb0")):
    raw = os.path.join(G, "%s_k_render_%s_raw.csv" % (tag, kern))
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    seg = int(open(os.path.join(G, "%s_k_render_%s_segments.txt" % (tag, kern))).read().split(":")[1])
    out = ["# ncu --set full --clock-control none --import-source on, %s<false>, one launch" % name,
           "# command: RTX_KERNEL=%s python bench.py --steps 1 --warmup 1 --spp 32 --no-cpu --no-count --no-rtow  (1200x800, reference mesh mix, %d segments)" % (kern, seg)]
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            out.append("%-82s %-16s %s" % (k, units[i], vals[i]))
    st = sorted(((float(vals[i]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for i, k in enumerate(hdr)
                 if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")), reverse=True)
    out.append("")
    out.append("# warps stalled per issue, by reason: " + "  ".join("%s %.2f" % (k, v) for v, k in st[:9]))
    src = os.path.join(G, "%s_k_render_%s_source.csv" % (tag, kern))
    sfx = "" if kern == "reg" else "_q"
    if kern == "reg":
        subprocess.check_call([sys.executable, os.path.join(P, "issue_json.py"), raw, str(seg), name,
                               "ncu --set full --clock-control none -k regex:k_render -s 1 -c 1 python bench.py --steps 1 --warmup 1 --spp 32 --no-cpu --no-count --no-rtow (tools/gpu_final.sh)",
                               os.path.join(P, "r02_k_render_issue.json")], stdout=subprocess.DEVNULL)
    open(os.path.join(P, "r02_k_render%s_ncu_summary.txt" % sfx), "w").write("\n".join(out) + "\n")
    for script, what, extra in (("block_budget.py", "block_budget", [str(seg), "40"]), ("sass_lines.py", "hot_lines", ["40"]), ("ncu_stalls.py", "stalls", [])):
        txt = subprocess.check_output([sys.executable, os.path.join(P, script), src, sass, mangled] + extra).decode()
        open(os.path.join(P, "r02_k_render%s_%s.txt" % (sfx, what)), "w").write(txt)
    print("\n".join(out))
t = {}
for r in csv.reader(open(os.path.join(G, tag + "_k_render_traffic.csv"))):
    if len(r) > 14 and r[12] in ("dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum", "gpu__time_duration.sum"):
        t[r[12]] = int(float(r[14]))
json.dump({"kernel": "k_render", "config": "1200x800, 500 spp, depth 50, mesh 9/6/3/8/6/3",
           "dram_bytes_read": t["dram__bytes_read.sum"], "dram_bytes_write": t["dram__bytes_write.sum"],
           "lts_t_bytes": t["lts__t_bytes.sum"], "l1tex_t_bytes": t["l1tex__t_bytes.sum"], "gpu_time_ns": t["gpu__time_duration.sum"],
           "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__t_bytes.sum,gpu__time_duration.sum --clock-control none (tools/gpu_final.sh)"},
          open(os.path.join(P, "r02_k_render_traffic.json"), "w"), indent=1)
for name in ("bench", "bench_reference", "bench_analytic", "bench_spp1", "bench_uhd64", "bench_stress", "bench_poolkernel", "bench_counted_analytic"):
    src = os.path.join(G, "%s_%s.json" % (tag, name))
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, "r02_%s.json" % name))
shutil.copy(os.path.join(G, tag + "_launches.csv"), os.path.join(P, "r02_launches.csv"))
print(json.dumps(t))
