#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "clustering" > gpurun_out/pytest_cluster.txt 2>&1
tail -30 gpurun_out/pytest_cluster.txt | cut -c1-250
timeout 300 python tools/rw_micro.py > gpurun_out/rw_micro.jsonl 2> gpurun_out/rw_micro.err; cat gpurun_out/rw_micro.jsonl; tail -2 gpurun_out/rw_micro.err
timeout 300 python bench.py --shape s3dis --points 1000000 --scenes 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s3dis_n1.json 2> gpurun_out/bench_s3dis_n1.err
python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_s3dis_n1.json') if l.startswith('{')][-1]);print('s3dis n1', d['value'], d['ms_per_step'], d['e2e']['value'])"
