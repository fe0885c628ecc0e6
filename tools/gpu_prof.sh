#!/bin/bash
# usage: tools/gpu_prof.sh <tag> <lib> <carveout> [spp]  -- one ncu --set full capture of k_render
tag=$1; lib=$2; carve=$3; spp=${4:-8}
export RTX_LIB=$PWD/rtxplay_b200/$lib RTX_CARVEOUT=$carve
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -o gpurun_out/prof_$tag -f python bench.py --steps 1 --warmup 1 --spp $spp --no-cpu --no-count > gpurun_out/ncu_full_$tag.log 2>&1
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/prof_${tag}_raw.csv')))
hdr,units,vals=rows[0],rows[1],rows[2]
print('PROF $tag ($lib carve $carve spp $spp)')
for k in ['gpu__time_duration.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct']:
    i=hdr.index(k); print('  %-66s %s %s'%(k,vals[i],units[i]))
PY
