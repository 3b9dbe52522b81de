#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 600 --timeout-method thread > gpurun_out/pytest_train.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_train.txt
tail -8 gpurun_out/pytest_train.txt | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nocpu.json 2> gpurun_out/bench.err
python -c "
import json;d=json.load(open('gpurun_out/bench_nocpu.json'));print({k:d[k] for k in ('value','ms_per_step')}, d['e2e'])"
timeout 900 python bench.py --mode train --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
python -c "
import json;d=json.load(open('gpurun_out/bench_train.json'));print({k:d[k] for k in ('value','ms_per_step','stages_ms')}, d['e2e']['value'])"
tail -2 gpurun_out/bench_train.err
