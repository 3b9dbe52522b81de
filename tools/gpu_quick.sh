#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 600 --timeout-method thread > gpurun_out/pytest_train.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_train.txt
tail -6 gpurun_out/pytest_train.txt | cut -c1-300
timeout 900 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
python -c "
import json;d=json.load(open('gpurun_out/bench_train.json'));print({k:d[k] for k in ('value','ms_per_step','ms_per_step_min_median_max','stages_ms')}, d['e2e']['value'], d['e2e']['ms_per_step_min_median_max'], d['parity'], d['cpu_baseline']['value'])"
tail -2 gpurun_out/bench_train.err
timeout 300 python tools/profile_train.py --rows 16 > gpurun_out/train_kernels.txt 2>&1
head -30 gpurun_out/train_kernels.txt | cut -c1-100,180-240
