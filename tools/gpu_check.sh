#!/bin/bash
# GPU check of the conv path: targeted parity tests first (under a timeout: a protocol bug in the persistent kernel
# would otherwise hang the box), then the conv microbench.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "tile_records or conv or smoke" > gpurun_out/pytest_conv.txt 2>&1
rc=$?
tail -15 gpurun_out/pytest_conv.txt
if [ $rc -ne 0 ]; then exit $rc; fi
timeout 300 python tools/conv_micro.py --shapes 32x32,64x64,128x128,6x32 --iters 5 > gpurun_out/micro_fp32.json 2> gpurun_out/micro_fp32.err
timeout 300 python tools/conv_micro.py --shapes 32x32,64x64,128x128 --precision bf16 --iters 5 > gpurun_out/micro_bf16.json 2> gpurun_out/micro_bf16.err
cat gpurun_out/micro_fp32.json gpurun_out/micro_bf16.json
tail -3 gpurun_out/micro_fp32.err
