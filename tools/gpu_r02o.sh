#!/bin/bash
mkdir -p gpurun_out
tools/gpu_sweep3.sh "librtx.so,RTX_KERNEL=reg librtx.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=75" 2>&1 | grep -E "SWEEP" | tee gpurun_out/r02o.log
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee -a gpurun_out/r02o.log
