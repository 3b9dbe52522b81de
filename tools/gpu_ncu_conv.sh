#!/bin/bash
# One ncu --set full capture (with source correlation) of the tensor-core conv kernel on the 32->32 microbench layer.
set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 2 -c 1 -f -o gpurun_out/conv_umma_v3 \
    python tools/conv_micro.py --shapes ${1:-32x32} --iters 1 --precision ${2:-fp32} > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
