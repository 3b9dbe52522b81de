#!/bin/bash
# One gpurun call: conv microbench, ncu --set full of the tensor-core conv kernel, launch list of a bench step.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
python tools/conv_micro.py --shapes 32x32,64x64,128x128 --iters 5 > gpurun_out/micro_fp32.json 2> gpurun_out/micro_fp32.err
python tools/conv_micro.py --shapes 32x32,64x64 --precision bf16 --iters 5 > gpurun_out/micro_bf16.json 2> gpurun_out/micro_bf16.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 2 -c 1 -f -o gpurun_out/conv_umma_c32 \
    python tools/conv_micro.py --shapes 32x32 --iters 1 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_v2.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 600 gpurun_out/bench.json
