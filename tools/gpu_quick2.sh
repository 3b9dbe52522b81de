#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "dense_rows or golden or full_size or batch_stream or drop_in" > gpurun_out/pytest_quick.txt 2>&1
tail -6 gpurun_out/pytest_quick.txt | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nocpu.json 2> gpurun_out/bench.err
python -c "
import json;d=json.load(open('gpurun_out/bench_nocpu.json'));print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','launches_total_per_step')}, d['e2e']['value'], d['e2e']['ms_per_step_min_median_max'])"
tail -2 gpurun_out/bench.err
