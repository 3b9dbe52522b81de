#!/bin/bash
# GPU check of the conv path: the targeted parity tests (every step under a timeout: a protocol bug in the persistent
# kernel traps through its barrier watchdog, and the timeouts bound what a hang can cost), then the conv microbench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 120 --timeout-method thread -k "tile_records or umma or morton or row_cache or neighbour or subm_conv_forward or fused_prologue or strided_and_inverse or dense_conv3d" > gpurun_out/pytest_conv.txt 2>&1
rc=$?
tail -25 gpurun_out/pytest_conv.txt
if [ $rc -ne 0 ]; then exit $rc; fi
timeout 200 python tools/conv_micro.py --shapes 32x32,64x32,6x32 --iters 5 > gpurun_out/micro_fp32.json 2> gpurun_out/micro_fp32.err
timeout 200 python tools/conv_micro.py --shapes 32x32,64x64 --precision bf16 --iters 5 > gpurun_out/micro_bf16.json 2> gpurun_out/micro_bf16.err
timeout 200 python tools/conv_micro.py --shapes 64x64,128x64 --level 2 --iters 5 > gpurun_out/micro_fp32_l2.json 2> gpurun_out/micro_l2.err
timeout 200 python tools/conv_micro.py --shapes 96x96,128x128 --level 3 --iters 5 > gpurun_out/micro_fp32_l3.json 2> gpurun_out/micro_l3.err
cat gpurun_out/micro_fp32.json gpurun_out/micro_bf16.json gpurun_out/micro_fp32_l2.json gpurun_out/micro_fp32_l3.json
tail -3 gpurun_out/micro_*.err
exit $rc
