#!/bin/bash
# usage (GPU box, repo root): tools/gpu_sanitize.sh  -- compute-sanitizer over tools/sanitize_driver.py, both path-tracing kernels
for kern in reg q; do
for tool in memcheck racecheck initcheck; do
  RTX_KERNEL=$kern timeout 400 compute-sanitizer --tool $tool python tools/sanitize_driver.py > gpurun_out/sanitize_${kern}_$tool.log 2>&1
  echo "$kern $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${kern}_$tool.log | tail -1)  [$(grep -c ' ok ' gpurun_out/sanitize_${kern}_$tool.log) scenes ok]"
done
done
