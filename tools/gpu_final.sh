#!/bin/bash
# usage (GPU box, repo root): tools/gpu_final.sh <tag>   -- the full evidence set of a round:
# GPU tests, both bench arms, the ncu launch list, one --set full capture of the render kernel (32 spp,
# both kernels) and a metrics-only pass at the bench configuration (DRAM / L2 / L1 bytes of one frame).
tag=${1:-rXX}
o=gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) | tee $o/${tag}_pytest.log
timeout 800 python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err; tail -c 300 $o/${tag}_bench.err
timeout 600 python bench.py --impl reference > $o/${tag}_bench_reference.json 2> $o/${tag}_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-count --no-rtow > $o/${tag}_launches.log 2>&1
for kern in reg q; do
  RTX_KERNEL=$kern timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -o $o/${tag}_k_render_$kern -f python bench.py --steps 1 --warmup 1 --spp 32 --no-cpu --no-count --no-rtow > $o/${tag}_ncu_full_$kern.log 2>&1
  ncu -i $o/${tag}_k_render_$kern.ncu-rep --page raw --csv > $o/${tag}_k_render_${kern}_raw.csv 2>/dev/null
  ncu -i $o/${tag}_k_render_$kern.ncu-rep --page source --csv > $o/${tag}_k_render_${kern}_source.csv 2>/dev/null
  grep -o '"segments_per_frame": [0-9]*' $o/${tag}_ncu_full_$kern.log | head -1 > $o/${tag}_k_render_${kern}_segments.txt
done
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__t_bytes.sum,gpu__time_duration.sum --clock-control none -k regex:k_render -s 1 -c 1 --csv --log-file $o/${tag}_k_render_traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-count --no-rtow > $o/${tag}_traffic.log 2>&1
python - <<PY
import json
for t in ("bench","bench_reference"):
    try:
        d=json.loads(open("$o/${tag}_%s.json"%t).read().strip().splitlines()[-1])
        print(t, "%.4g %s  %.1f ms/step"%(d["value"],d["unit"],d["ms_per_step"]), d.get("e2e"), d["config"].get("build_ms"))
    except Exception as e: print(t,"FAILED",e)
PY
tail -3 $o/${tag}_k_render_traffic.csv
# the pool kernel on the bench configuration
RTX_KERNEL=q timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-rtow > $o/${tag}_bench_poolkernel.json 2> $o/${tag}_bench_poolkernel.err
# the other configurations of BASELINE.json, short runs without the CPU / counted legs
for cfg in "analytic:--mode analytic" "spp1:--spp 1 --steps 20" "uhd64:--width 3840 --height 2160 --spp 64" "stress:--scene grid316 --spp 64" "counted_analytic:--mode analytic --steps 1 --warmup 1" "poolkernel:"; do
  name=${cfg%%:*}; opts=${cfg#*:}
  extra="--no-cpu --no-count --no-rtow"; [ "$name" = counted_analytic ] && extra="--no-cpu --no-rtow"
  [ "$name" = poolkernel ] || timeout 600 python bench.py $extra --steps 3 --warmup 3 $opts > $o/${tag}_bench_$name.json 2> $o/${tag}_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("$o/${tag}_bench_$name.json").read().strip().splitlines()[-1])
    print("EXTRA $name: %.3f Gseg/s  %.2f ms/frame  build %s" % (d["value"]/1e9, d["ms_per_step"], {k: round(v,2) for k,v in d["config"]["build_ms"].items() if not isinstance(v, dict)}))
except Exception as e: print("EXTRA $name FAILED", e)
PY
done
