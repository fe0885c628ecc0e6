#!/bin/bash
# usage (here): make -C rtxplay_b200/csrc trim ; gpurun -- tools/gpu_trim.sh
# times the instruction-trimming builds against the shipped one and runs the GPU parity tests on the combined build
tools/gpu_sweep2.sh "librtx.so:35:20 librtx_fp.so:35:20 librtx_lp.so:35:20 librtx_f2.so:35:20 librtx_trim.so:35:20 librtx.so:35:20" 2>&1 | grep SWEEP | tee gpurun_out/sweep_trim.txt
RTX_LIB=$PWD/rtxplay_b200/librtx_trim.so timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_trim.log
