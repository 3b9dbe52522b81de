#!/bin/bash
# 8-GPU pass: data-parallel training step (configs[2]) and S3DIS rooms sharded over 4 / 8 GPUs (configs[3]).
mkdir -p gpurun_out
run() { N=$1; PORT=$2; shift 2; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N "$@"; }
run 8 29500 --mode train --steps 6 --warmup 3 > gpurun_out/bench_train_n8.json 2> gpurun_out/bench_train_n8.err
tail -c 700 gpurun_out/bench_train_n8.json; tail -2 gpurun_out/bench_train_n8.err
run 8 29501 --shape s3dis --points 1000000 --scenes 1 --steps 10 --warmup 3 > gpurun_out/bench_s3dis_n8.json 2> gpurun_out/bench_s3dis_n8.err
python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_s3dis_n8.json') if l.startswith('{')][-1]);print('s3dis n8', d['value'], d['ms_per_step'], d['e2e']['value'])"
run 4 29502 --shape s3dis --points 1000000 --scenes 1 --steps 10 --warmup 3 > gpurun_out/bench_s3dis_n4.json 2> gpurun_out/bench_s3dis_n4.err
python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_s3dis_n4.json') if l.startswith('{')][-1]);print('s3dis n4', d['value'], d['ms_per_step'], d['e2e']['value'])"
