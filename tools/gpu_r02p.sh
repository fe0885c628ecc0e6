#!/bin/bash
mkdir -p gpurun_out
tools/gpu_sweep3.sh "librtx.so,RTX_KERNEL=reg librtx_swap.so,RTX_KERNEL=reg librtx_c19.so,RTX_KERNEL=reg librtx_c20.so,RTX_KERNEL=reg librtx_c16.so,RTX_KERNEL=reg librtx.so,RTX_KERNEL=reg,RTX_CARVEOUT=30 librtx.so,RTX_KERNEL=reg,RTX_CARVEOUT=42 librtx.so,RTX_KERNEL=reg" 2>&1 | grep -E "SWEEP" | tee gpurun_out/r02p.log
