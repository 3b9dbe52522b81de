#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/sync_bn_check.py > gpurun_out/sync_bn_check.txt 2>&1
grep -E "^\{|Error|assert" gpurun_out/sync_bn_check.txt | head -5 | cut -c1-1200
for T in peer nccl; do
WSIS_SYNC_BN_TRANSPORT=$T timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --mode train --steps 10 --warmup 3 > gpurun_out/bench_train_n2_$T.json 2> gpurun_out/bench_train_n2_$T.err
python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_train_n2_$T.json') if l.startswith('{')][-1]);print('$T', d['value'], d['ms_per_step_min_median_max'], d['stages_ms'], d['e2e']['value'], d['e2e']['ms_per_step_min_median_max'])"
done
timeout 300 python bench.py --mode train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_n1b.json 2> gpurun_out/bench_train_n1b.err
python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_train_n1b.json') if l.startswith('{')][-1]);print('n1', d['value'], d['ms_per_step_min_median_max'], d['stages_ms'], d['e2e']['value'], d['e2e']['ms_per_step_min_median_max'])"
