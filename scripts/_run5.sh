set -x
python -m pytest tests/test_glue_gpu.py -m gpu -x -q 2>&1 | tail -15
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python scripts/time_train.py --profile > gpurun_out/r02_train_profile_b.txt 2>&1; tail -3 gpurun_out/r02_train_profile_b.txt
python scripts/time_train.py --graph 2>&1 | tail -2
python scripts/bench_hbm_kernels.py 2>&1 | head -1
