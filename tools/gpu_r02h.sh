#!/bin/bash
mkdir -p gpurun_out
tools/gpu_sweep3.sh "librtx.so,RTX_KERNEL=reg librtx_noinl.so,RTX_KERNEL=reg librtx_q64.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=75 librtx_q64.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=60 librtx_q56.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=75 librtx_q56.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=60 librtx_q72.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=75 librtx_noinl.so,RTX_KERNEL=reg librtx.so,RTX_KERNEL=reg" 2>&1 | grep -E "SWEEP|pool kernel" | tee gpurun_out/r02h.log
RTX_KERNEL=reg timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee -a gpurun_out/r02h.log
