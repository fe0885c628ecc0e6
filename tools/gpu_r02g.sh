#!/bin/bash
mkdir -p gpurun_out
tools/gpu_sweep3.sh "librtx_q64s13.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=75 librtx_q64s13.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=85 librtx_q64s11.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=75 librtx_q64s11.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=85 librtx_q64s11np.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=75 librtx_q80s11.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=75 librtx_q80s11.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=85" 2>&1 | grep -E "SWEEP|pool kernel" | tee gpurun_out/r02g.log
