#!/bin/bash
# usage: tools/gpu_build_sweep.sh "<lib> <lib> ..." [n_spheres]   -- stage times of one big LBVH build per library (tools/build_scale.py)
for lib in $1; do
  RTX_LIB=$PWD/rtxplay_b200/$lib timeout 300 python tools/build_scale.py ${2:-6500} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read().split('BUILD ')[1]); print('BUILD $lib', {k:round(v['ms'],3) for k,v in d['stages'].items()})"
done
