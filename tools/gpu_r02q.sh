#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r02q.log
timeout 300 python tools/build_scale.py 64 --render 2>&1 | grep -E "BUILD|rror" | tee -a gpurun_out/r02q.log
timeout 600 python tools/build_scale.py 6500 --render 2>&1 | grep -E "BUILD|rror" | tee -a gpurun_out/r02q.log
RTX_NO_COOP=1 timeout 300 python tools/build_scale.py 64 2>&1 | grep -E "BUILD|rror" | tee -a gpurun_out/r02q.log
python bench.py --steps 2 --warmup 1 --no-cpu --no-count --no-rtow 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('BENCH %.1f ms/frame'%d['ms_per_step'], d['config']['build_ms'])" | tee -a gpurun_out/r02q.log
