#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread > gpurun_out/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -15 gpurun_out/pytest_gpu.txt | cut -c1-300
timeout 900 python bench.py --mode train --steps 6 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
tail -c 1800 gpurun_out/bench_train.json; tail -3 gpurun_out/bench_train.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','launches_total_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['all_sparse_conv'])"
tail -3 gpurun_out/bench.err
python tools/grad_conditioning.py 9000 > gpurun_out/grad_conditioning.txt 2>&1; tail -11 gpurun_out/grad_conditioning.txt | cut -c1-200
