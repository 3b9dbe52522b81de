#!/bin/bash
# N-GPU pass (default 2): data-parallel training step (SyncBN + NCCL gradient all-reduce), S3DIS rooms sharded over the
# GPUs, inference.
N=${1:-2}
mkdir -p gpurun_out
TR="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29500 bench.py --gpus $N --mode train --steps 6 --warmup 3 > gpurun_out/bench_train_n$N.json 2> gpurun_out/bench_train_n$N.err
tail -c 1500 gpurun_out/bench_train_n$N.json; tail -3 gpurun_out/bench_train_n$N.err
$TR --master-port 29501 bench.py --gpus $N --shape s3dis --points 1000000 --scenes 1 --steps 10 --warmup 3 > gpurun_out/bench_s3dis_n$N.json 2> gpurun_out/bench_s3dis_n$N.err
tail -c 1500 gpurun_out/bench_s3dis_n$N.json; tail -3 gpurun_out/bench_s3dis_n$N.err
$TR --master-port 29502 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 800 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
