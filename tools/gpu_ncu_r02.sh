#!/bin/bash
# ncu --set full captures (one launch each) of the round-2 tensor-core kernels + the launch list of a bench step.
mkdir -p gpurun_out
NCU="timeout 400 ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:conv_umma -s 2 -c 1 -o gpurun_out/r02_conv_umma_l2 python tools/conv_micro.py --level 2 --shapes 64x64 --iters 1 > gpurun_out/ncu_conv.log 2>&1
$NCU -k regex:wgrad_umma -s 2 -c 1 -o gpurun_out/r02_wgrad_umma_l2 python tools/wgrad_micro.py --level 2 --shapes 64x64 --iters 1 > gpurun_out/ncu_wgrad.log 2>&1
$NCU -k regex:ecc_messages -s 2 -c 1 -o gpurun_out/r02_ecc_messages python tools/ecc_micro.py 1 > gpurun_out/ncu_ecc.log 2>&1
$NCU -k regex:mlp_head -s 1 -c 1 -o gpurun_out/r02_mlp_head python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_head.log 2>&1
for f in r02_conv_umma_l2 r02_wgrad_umma_l2 r02_ecc_messages r02_mlp_head; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r02_final.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python tools/ecc_micro.py > gpurun_out/ecc_micro.jsonl 2>&1; cat gpurun_out/ecc_micro.jsonl | tail -2
python tools/cluster_micro.py > gpurun_out/cluster_micro.jsonl 2>&1; tail -1 gpurun_out/cluster_micro.jsonl
ls -la gpurun_out/*.ncu-rep | tail -5
