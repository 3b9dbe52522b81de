#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 600 --timeout-method thread > gpurun_out/pytest_train.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_train.txt
tail -12 gpurun_out/pytest_train.txt | cut -c1-300
python tools/wgrad_micro.py --shapes 32x32,64x64 --debug 0 2>&1 | grep fp32
timeout 600 python tools/profile_train.py --rows 22 > gpurun_out/train_kernels.txt 2>&1
head -40 gpurun_out/train_kernels.txt | cut -c1-100,180-240
