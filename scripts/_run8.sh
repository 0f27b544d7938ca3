python scripts/time_train.py --profile 2>&1 | cut -c1-100,165- | grep -v "^-" | grep "wgrad"
python scripts/time_train.py --graph 2>&1 | tail -1
CSBSR_WGRAD_ITEMS_PER_SM=2 python scripts/time_train.py --graph 2>&1 | tail -1
python -m pytest tests/test_train_gpu.py -m gpu -x -q 2>&1 | tail -2
