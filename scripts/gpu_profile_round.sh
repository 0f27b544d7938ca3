#!/bin/bash
# One GPU session that produces everything profiles/ summarises (run under gpurun; outputs into gpurun_out/).
set -x
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline --train-steps 1 --no-graph > gpurun_out/launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 40 -c 4 -o gpurun_out/prof_wgrad_r01 -f \
    python scripts/time_train.py --steps 1 --warmup 1 > gpurun_out/prof_wgrad.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 700 -c 8 -o gpurun_out/prof_conv_r01b -f \
    python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline --no-train --no-graph > gpurun_out/prof_conv.log 2>&1
ls -la gpurun_out/
