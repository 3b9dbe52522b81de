#!/bin/bash
# GPU pass for the training step: full GPU suite (the conv / module code paths changed), the training bench at N=1,
# the configs[4] micro sweep.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread -x > gpurun_out/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -25 gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --mode train --steps 6 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
tail -c 3500 gpurun_out/bench_train.json; tail -5 gpurun_out/bench_train.err
timeout 600 python tools/conv_micro.py --sweep --iters 5 > gpurun_out/sweep_fp32.jsonl 2> gpurun_out/sweep.err
tail -5 gpurun_out/sweep_fp32.jsonl; tail -3 gpurun_out/sweep.err
