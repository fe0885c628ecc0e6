#!/bin/bash
# usage: tools/gpu_sweep3.sh "<lib>[,VAR=val]... ..." [extra bench args]   -- one short bench run per configuration
# e.g.   tools/gpu_sweep3.sh "librtx.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=60 librtx.so,RTX_KERNEL=reg"
for cfg in $1; do
  IFS=, read -ra parts <<< "$cfg"
  lib=${parts[0]}
  envs=("${parts[@]:1}")
  env RTX_VERBOSE=1 RTX_LIB=$PWD/rtxplay_b200/$lib "${envs[@]}" timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-count ${2:-} > gpurun_out/sweep_tmp.json 2> gpurun_out/sweep_tmp.err
  grep -E "rtx_init|rtx:" gpurun_out/sweep_tmp.err | sort -u | head -4
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/sweep_tmp.json').read().strip().splitlines()[-1])
    print('SWEEP $cfg: %.3f Gseg/s  %.1f ms/frame' % (d['value']/1e9, d['ms_per_step']))
except Exception as e:
    print('SWEEP $cfg FAILED', e, open('gpurun_out/sweep_tmp.err').read()[-400:])
PY
done
