#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 200 -k "ecc or golden or full_size or mlp_head" > gpurun_out/pytest_ecc.txt 2>&1
tail -12 gpurun_out/pytest_ecc.txt | cut -c1-250
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','launches_total_per_step')}, d['e2e']['value'], d['roofline']['frac']);print(json.dumps(d['roofline']['by_layer_shape']))"
tail -2 gpurun_out/bench.err
timeout 200 python tools/profile_train.py --mode infer --rows 30 > gpurun_out/infer_kernels.txt 2>&1
