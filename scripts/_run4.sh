set -x
python -m pytest tests/test_metrics_gpu.py tests/test_train_gpu.py -m gpu -x -q -k "degrade or philox or b8" 2>&1 | tail -5
python scripts/bench_hbm_kernels.py 2>&1 | head -2
python scripts/time_train.py --profile > gpurun_out/r02_train_profile_a.txt 2>&1; tail -3 gpurun_out/r02_train_profile_a.txt
