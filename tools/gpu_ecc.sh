#!/bin/bash
# ECC-GRU fused kernel: parity tests (module-level and the golden network), then the bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout 120 --timeout-method thread -k "ecc or golden or drop_in or s3dis" > gpurun_out/pytest_ecc.txt 2>&1
tail -25 gpurun_out/pytest_ecc.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ecc.json 2> gpurun_out/bench_ecc.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_ecc.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','ms_per_step_min_median_max','gpu_launches')}); print(d['e2e'])"
tail -3 gpurun_out/bench_ecc.err
