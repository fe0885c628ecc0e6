#!/bin/bash
mkdir -p gpurun_out
for spp in 16 32 63 125 250 500; do
python bench.py --steps 4 --warmup 2 --spp $spp --no-cpu --no-count --no-rtow --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('SPP $spp: %.3f ms/frame  %.3f Gseg/s'%(d['ms_per_step'], d['value']/1e9))" | tee -a gpurun_out/r02r.log
done
