#!/bin/bash
# usage: tools/gpu_sweep.sh "<lib>:<carveout> ..."   -- short bench of experimental builds
for cfg in $1; do
  lib=${cfg%%:*}; carve=${cfg##*:}
  RTX_LIB=$PWD/rtxplay_b200/$lib RTX_CARVEOUT=$carve timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-count ${2:-} > gpurun_out/sweep_tmp.json 2> gpurun_out/sweep_tmp.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/sweep_tmp.json').read().strip().splitlines()[-1])
    print('SWEEP $lib carve=$carve: %.3f Gseg/s  %.1f ms/frame' % (d['value']/1e9, d['ms_per_step']))
except Exception as e:
    print('SWEEP $lib carve=$carve FAILED', e, open('gpurun_out/sweep_tmp.err').read()[-300:])
PY
done
