#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -k "batch_stream" > gpurun_out/pytest_stream.txt 2>&1
tail -3 gpurun_out/pytest_stream.txt | cut -c1-300
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --stream-variants > gpurun_out/bench_stream.json 2> gpurun_out/bench_stream.err
python -c "
import json;d=json.load(open('gpurun_out/bench_stream.json'));print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e']['ms_per_step_min_median_max'], d.get('stream_variants'))"
grep -E "Error|error" gpurun_out/bench_stream.err | head -3
