#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_multi.py 2>&1 | grep -v "arena\|rtx_init:" | tail -6 | tee gpurun_out/r02l.log
timeout 600 python -m pytest tests -m gpu -q -x -k "multi_device or rtwo" --durations=5 2>&1 | tail -15 | tee -a gpurun_out/r02l.log
