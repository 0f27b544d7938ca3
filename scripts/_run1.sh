set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err; tail -c 600 gpurun_out/bench_ref_r02.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_ncu.log 2>&1; tail -2 gpurun_out/smoke_ncu.log
SAN_TIMEOUT=300 bash scripts/sanitize.sh 2>&1 | tail -12
