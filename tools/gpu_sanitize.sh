#!/bin/bash
# usage (GPU box, repo root): tools/gpu_sanitize.sh  -- compute-sanitizer over tools/sanitize_driver.py
for tool in memcheck racecheck initcheck; do
  timeout 240 compute-sanitizer --tool $tool python tools/sanitize_driver.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)  [$(grep -c ' ok ' gpurun_out/sanitize_$tool.log) scenes ok]"
done
