#!/bin/bash
# round 2, session e: pool kernel v2 (packed queues, lean stack, inline shading)
mkdir -p gpurun_out
echo "== GPU tests with the pool kernel" | tee gpurun_out/r02e.log
RTX_KERNEL=q timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee -a gpurun_out/r02e.log
echo "== sweep" | tee -a gpurun_out/r02e.log
tools/gpu_sweep3.sh "librtx.so,RTX_KERNEL=reg librtx.so,RTX_KERNEL=q librtx_q64s11.so,RTX_KERNEL=q librtx_q64s9.so,RTX_KERNEL=q librtx_q64s9.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=75 librtx_q64s9c16.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=75 librtx_q48s9.so,RTX_KERNEL=q librtx_q48s9.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=75 librtx_q48s9c20.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=60 librtx_q48s9c20.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=75 librtx_q48s9c20.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=85 librtx_q48s11.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=75" 2>&1 | grep -E "SWEEP|pool kernel" | tee -a gpurun_out/r02e.log
