set -x
python -m pytest tests/test_glue_gpu.py tests/test_train_gpu.py -m gpu -x -q 2>&1 | tail -8
python scripts/time_train.py --layers 2>&1 | grep -A12 "== wgrad\|== conv"
python scripts/time_train.py --graph 2>&1 | tail -1
