#!/bin/bash
# round 2: multi-GPU on real peers (run with gpurun --gpus 8)
mkdir -p gpurun_out
nvidia-smi -L | head -8 | tee gpurun_out/r02n_multi8.log
RTX_VERBOSE=1 timeout 600 python tools/multi_check.py 8 --c4 2>&1 | grep -E "MULTI|rtx_init_multi|Error|error" | tee -a gpurun_out/r02n_multi8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --width 3840 --height 2160 --spp 4096 --steps 2 --warmup 1 --no-e2e --no-cpu --no-count > gpurun_out/r02n_bench_c4_n8.json 2> gpurun_out/r02n_bench_c4_n8.err
tail -c 300 gpurun_out/r02n_bench_c4_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02n_bench_c3_n8.json 2> gpurun_out/r02n_bench_c3_n8.err
python - <<PY
import json
for t in ("c4_n8","c3_n8"):
    try:
        d=json.loads(open("gpurun_out/r02n_bench_%s.json"%t).read().strip().splitlines()[-1])
        print("BENCH", t, "%.3f Gseg/s %.1f ms/frame"%(d["value"]/1e9, d["ms_per_step"]), d["e2e"].get("ms_per_frame"))
    except Exception as e: print(t,"FAILED",e)
PY
