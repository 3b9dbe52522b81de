#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --geometry-prefetch > gpurun_out/bench_prefetch_$i.json 2> gpurun_out/bench_prefetch.err
python -c "
import json;d=json.load(open('gpurun_out/bench_prefetch_$i.json'));print($i, round(d['value'],1), d['ms_per_step_min_median_max'], round(d['e2e']['value'],1), d['e2e']['ms_per_step_min_median_max'], round(d['single_stream']['value'],1))"
done
grep -E "Error" gpurun_out/bench_prefetch.err | head -2
