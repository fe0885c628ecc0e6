#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee gpurun_out/r02k.log
RTX_KERNEL=q timeout 900 python -m pytest tests -m gpu -q -k "multi_device or counted or frame_equals" 2>&1 | tail -5 | tee -a gpurun_out/r02k.log
