#!/bin/bash
# round 2, session a: time the instruction-trimming builds and name the GPU test each one breaks (if any)
mkdir -p gpurun_out
tools/gpu_sweep2.sh "librtx.so:35:20 librtx_fp.so:35:20 librtx_lp.so:35:20 librtx_f2.so:35:20 librtx_trim.so:35:20 librtx.so:35:20" 2>&1 | grep SWEEP | tee gpurun_out/r02a_sweep_trim.txt
for v in fp lp f2 trim; do
  echo "== tests with librtx_$v.so" | tee -a gpurun_out/r02a_pytest_trim.log
  RTX_LIB=$PWD/rtxplay_b200/librtx_$v.so timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee -a gpurun_out/r02a_pytest_trim.log
done
