set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python scripts/time_train.py --profile > gpurun_out/r02_train_profile_c.txt 2>&1; tail -3 gpurun_out/r02_train_profile_c.txt
python scripts/time_train.py --graph 2>&1 | tail -2
python bench.py --no-cpu-baseline > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; tail -c 1500 gpurun_out/bench_r02c.json; tail -3 gpurun_out/bench_r02c.err
