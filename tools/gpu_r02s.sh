#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r02s.log
for lib in librtx.so librtx_u32.so librtx_u16.so librtx_b00.so librtx_t1.so; do
for spp in 63 500; do
RTX_LIB=$PWD/rtxplay_b200/$lib python bench.py --steps 4 --warmup 2 --spp $spp --no-cpu --no-count --no-rtow --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib SPP $spp: %.3f ms/frame  %.3f Gseg/s'%(d['ms_per_step'], d['value']/1e9))" | tee -a gpurun_out/r02s.log
done; done
