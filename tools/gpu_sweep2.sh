#!/bin/bash
# usage: sweep2.sh "<lib>:<carve>:<ctas> ..."
for cfg in $1; do
  IFS=: read lib carve ctas <<< "$cfg"
  RTX_VERBOSE=1 RTX_CTAS_PER_SM=$ctas RTX_LIB=$PWD/rtxplay_b200/$lib RTX_CARVEOUT=$carve timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-count ${2:-} > gpurun_out/sweep_tmp.json 2> gpurun_out/sweep_tmp.err
  grep -m1 "rtx_init" gpurun_out/sweep_tmp.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/sweep_tmp.json').read().strip().splitlines()[-1])
    print('SWEEP $lib carve=$carve ctas=$ctas: %.3f Gseg/s  %.1f ms/frame' % (d['value']/1e9, d['ms_per_step']))
except Exception as e:
    print('SWEEP $lib carve=$carve FAILED', e, open('gpurun_out/sweep_tmp.err').read()[-300:])
PY
done
