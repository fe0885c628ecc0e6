#!/bin/bash
# round 2, session d: ncu capture of the pool kernel
mkdir -p gpurun_out
export RTX_KERNEL=q
tag=r02d_q; lib=${1:-librtx_q64s9.so}; spp=${2:-8}
export RTX_LIB=$PWD/rtxplay_b200/$lib
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -o gpurun_out/prof_$tag -f python bench.py --steps 1 --warmup 1 --spp $spp --no-cpu --no-count > gpurun_out/ncu_full_$tag.log 2>&1
tail -3 gpurun_out/ncu_full_$tag.log
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv > gpurun_out/prof_${tag}_source.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/prof_${tag}_raw.csv')))
hdr,units,vals=rows[0],rows[1],rows[2]
for k in ['gpu__time_duration.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed']:
    try:
        i=hdr.index(k); print('%-66s %s %s'%(k,vals[i],units[i]))
    except ValueError: print('missing',k)
PY
