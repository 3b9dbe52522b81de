#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus $N --mode train --steps 10 --warmup 3 > gpurun_out/bench_train_n$N.json 2> gpurun_out/bench_train_n$N.err
python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_train_n$N.json') if l.startswith('{')][-1]);print('train n$N', d['value'], d['ms_per_step_min_median_max'], d['stages_ms'], d['e2e']['value'], d['e2e']['ms_per_step_min_median_max'])"
tail -2 gpurun_out/bench_train_n$N.err | cut -c1-300
