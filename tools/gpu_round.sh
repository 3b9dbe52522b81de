#!/bin/bash
# Full GPU pass of a round: every GPU parity test, smoke, the bench (with the CPU baseline), the full-size parity log.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread > gpurun_out/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -8 gpurun_out/pytest_gpu.txt | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "full_size and reference" > gpurun_out/parity_full_size.txt 2>&1; grep -E "bf16 end-to-end|passed|failed" gpurun_out/parity_full_size.txt | cut -c1-400
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','launches_total_per_step')}, d['e2e'], d['roofline']['frac'], d['cpu_baseline']['value'], d['parity']['max_rel'], d['parity']['rulebooks_equal'])"
tail -3 gpurun_out/bench.err
