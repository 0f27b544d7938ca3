#!/bin/bash
# One GPU session that produces everything profiles/ summarises (run under gpurun; outputs into gpurun_out/).  TAG=r02 by default.
set -x
TAG=${TAG:-r02}
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
timeout 300 python scripts/profile_convs.py > gpurun_out/${TAG}_conv_layers.txt 2>&1
timeout 300 python scripts/profile_eval_step.py > gpurun_out/${TAG}_eval_step_kernels.txt 2>&1
timeout 300 python scripts/time_train.py --profile --layers > gpurun_out/${TAG}_train_step_kernels.txt 2>&1
timeout 300 python scripts/bench_hbm_kernels.py > gpurun_out/${TAG}_hbm_kernels.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline --train-steps 1 --no-graph > gpurun_out/launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_kernel -s 40 -c 4 -o gpurun_out/prof_wgrad_${TAG} -f \
    python scripts/time_train.py --steps 1 --warmup 1 > gpurun_out/prof_wgrad.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 700 -c 8 -o gpurun_out/prof_conv_${TAG} -f \
    python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline --no-train --no-graph > gpurun_out/prof_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -c 4 -o gpurun_out/prof_chain_${TAG} -f \
    python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline --no-train --no-graph > gpurun_out/prof_chain.log 2>&1
ls -la gpurun_out/ | tail -20
