#!/bin/bash
# One GPU session that produces everything profiles/ summarises (run under gpurun; outputs into gpurun_out/).  TAG=r02 by default.
set -x
TAG=${TAG:-r02}
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
timeout 300 python scripts/profile_convs.py 16 > gpurun_out/${TAG}_conv_layers.txt 2>&1
timeout 300 python scripts/profile_eval_step.py > gpurun_out/${TAG}_eval_step_kernels.txt 2>&1
timeout 300 python scripts/time_train.py --profile --layers > gpurun_out/${TAG}_train_step_kernels.txt 2>&1
timeout 300 python scripts/bench_hbm_kernels.py > gpurun_out/${TAG}_hbm_kernels.txt 2>&1
# launch lists (time + DRAM bytes per launch) of exactly one eval pass (16 images) and one training step
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_eval_${TAG}.csv python scripts/ncu_one_pass.py eval 16 > gpurun_out/launches_eval.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_train_${TAG}.csv python scripts/ncu_one_pass.py train > gpurun_out/launches_train.log 2>&1
# --set full captures: 40 consecutive conv launches of the first KBPN stages, the fused chains, the weight-gradient kernel
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_igemm -s 20 -c 40 -o /tmp/prof_conv_${TAG} -f \
    python scripts/ncu_one_pass.py eval 16 > gpurun_out/prof_conv.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:chain_kernel -c 2 -o /tmp/prof_chain_${TAG} -f \
    python scripts/ncu_one_pass.py eval 16 > gpurun_out/prof_chain.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:hd_fused -c 1 -o /tmp/prof_hd_${TAG} -f \
    python scripts/ncu_one_pass.py eval 16 > gpurun_out/prof_hd.log 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_wgrad_kernel -c 40 -o /tmp/prof_wgrad_${TAG} -f \
    python scripts/ncu_one_pass.py train > gpurun_out/prof_wgrad.log 2>&1
# the reports stay on the box (40 launches x 2.5 MB exceed the 64 MiB that travel back): raw metric pages as CSV instead
for k in conv chain hd wgrad; do
    ncu -i /tmp/prof_${k}_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${k}_${TAG}_raw.csv 2> /dev/null
done
ls -la gpurun_out/ | tail -24
