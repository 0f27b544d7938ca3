set -x
python -m pytest tests/test_metrics_gpu.py tests/test_train_gpu.py -m gpu -x -q -k "degrade or philox or b8" 2>&1 | tail -15
python scripts/bench_hbm_kernels.py > gpurun_out/r02_hbm_kernels.txt 2>&1; cat gpurun_out/r02_hbm_kernels.txt | head -14
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_ncu.log 2>&1; tail -2 gpurun_out/smoke_ncu.log
grep -c "conv_wgrad" gpurun_out/smoke_launches.csv
