#!/bin/bash
o=gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-count --no-rtow --no-e2e > $o/chk_bench.json 2> $o/chk_bench.err
python - <<PY
import json
d=json.loads(open("$o/chk_bench.json").read().strip().splitlines()[-1]); print("CHK %.3f Gseg/s %.2f ms"%(d["value"]/1e9,d["ms_per_step"]))
PY
