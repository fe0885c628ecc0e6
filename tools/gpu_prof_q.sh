#!/bin/bash
# usage: tools/gpu_prof_q.sh <tag> <lib> <q carveout> [spp] [kernel]  -- one ncu --set full capture of the render kernel
tag=$1; lib=$2; carve=$3; spp=${4:-32}; kern=${5:-q}
export RTX_KERNEL=$kern RTX_Q_CARVEOUT=$carve RTX_LIB=$PWD/rtxplay_b200/$lib
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -o gpurun_out/prof_$tag -f python bench.py --steps 1 --warmup 1 --spp $spp --no-cpu --no-count > gpurun_out/ncu_full_$tag.log 2>&1
grep -o '"segments_per_frame": [0-9]*' gpurun_out/ncu_full_$tag.log | head -1
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv > gpurun_out/prof_${tag}_source.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/prof_${tag}_raw.csv')))
hdr,units,vals=rows[0],rows[1],rows[2]
print('PROF $tag ($lib carve $carve spp $spp kernel $kern)')
for k in ['gpu__time_duration.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_wait_per_warp_active.pct','smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct','smsp__warp_issue_stalled_no_instruction_per_warp_active.pct','smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct','smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct','smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct','smsp__warp_issue_stalled_barrier_per_warp_active.pct','smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct','smsp__warp_issue_stalled_not_selected_per_warp_active.pct']:
    try:
        i=hdr.index(k); print('  %-78s %s %s'%(k,vals[i],units[i]))
    except ValueError: pass
PY
