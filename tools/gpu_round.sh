#!/bin/bash
# Full GPU pass of a round: every GPU parity test, smoke, the conv microbench, one ncu --set full capture of the
# dominant kernel, the launch list of a bench step and the bench itself (with the CPU baseline).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -5 gpurun_out/pytest_gpu.txt
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
timeout 200 python tools/conv_micro.py --shapes 32x32,64x64,128x128,6x32 --iters 5 > gpurun_out/micro_fp32.json 2> gpurun_out/micro_fp32.err
timeout 200 python tools/conv_micro.py --shapes 32x32,64x64,128x128 --precision bf16 --iters 5 > gpurun_out/micro_bf16.json 2> gpurun_out/micro_bf16.err
timeout 200 python tools/conv_micro.py --shapes 64x64,96x96 --level 2 --iters 5 > gpurun_out/micro_fp32_l2.json 2> gpurun_out/micro_l2.err
cat gpurun_out/micro_fp32.json gpurun_out/micro_bf16.json gpurun_out/micro_fp32_l2.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 2 -c 1 -f -o gpurun_out/conv_umma_v3 \
    python tools/conv_micro.py --level 2 --shapes 64x64 --iters 1 > gpurun_out/ncu_full.log 2>&1
if [ "$1" != "nolist" ]; then
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 700 -c 640 --csv --log-file gpurun_out/launches_v3.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
fi
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
tail -c 600 gpurun_out/bench_bf16.json
