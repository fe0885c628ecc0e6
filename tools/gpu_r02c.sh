#!/bin/bash
# round 2, session c: first run of the compacting-pool kernel (k_render_q)
mkdir -p gpurun_out
echo "== smoke with the pool kernel" | tee gpurun_out/r02c.log
RTX_KERNEL=q RTX_VERBOSE=1 timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee -a gpurun_out/r02c.log
echo "== GPU tests with the pool kernel" | tee -a gpurun_out/r02c.log
RTX_KERNEL=q timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee -a gpurun_out/r02c.log
echo "== sweep" | tee -a gpurun_out/r02c.log
tools/gpu_sweep3.sh "librtx.so,RTX_KERNEL=reg librtx.so,RTX_KERNEL=q librtx.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=50 librtx.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=70 librtx_q64.so,RTX_KERNEL=q librtx_q64.so,RTX_KERNEL=q,RTX_Q_CARVEOUT=45 librtx_q128.so,RTX_KERNEL=q librtx_q96s9.so,RTX_KERNEL=q librtx_q64s9.so,RTX_KERNEL=q" 2>&1 | grep -E "SWEEP|rtx_init" | tee -a gpurun_out/r02c.log
