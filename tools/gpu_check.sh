#!/bin/bash
# usage (on the GPU box, from the repo root): tools/gpu_check.sh <tag> [test|notest] [prof|noprof]
# runs the GPU parity tests, the default bench, and one ncu --set full capture of k_render
tag=${1:-x}
if [ "${2:-test}" = "test" ]; then
  (timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) | tee gpurun_out/pytest_$tag.log
fi
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -c 400 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$tag.json').read().strip().splitlines()[-1])
print('BENCH $tag: %.3f Gseg/s  %.1f ms/frame  e2e %.1f ms  kernel %.1f ms  launches %d  clocks %s' % (d['value']/1e9, d['ms_per_step'], d['e2e']['ms_per_frame'], d['roofline']['kernel_ms'], d['gpu_launches'], d['clocks']))
PY
if [ "${3:-prof}" = "prof" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -o gpurun_out/prof_$tag -f python bench.py --steps 1 --warmup 1 --spp 8 --no-cpu --no-count > gpurun_out/ncu_full_$tag.log 2>&1
  ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
  python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/prof_${tag}_raw.csv')))
hdr,units,vals=rows[0],rows[1],rows[2]
for k in ['gpu__time_duration.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct']:
    i=hdr.index(k); print('%-66s %s %s'%(k,vals[i],units[i]))
PY
fi
