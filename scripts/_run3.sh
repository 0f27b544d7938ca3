set -x
python -m pytest tests/test_metrics_gpu.py tests/test_train_gpu.py -m gpu -x -q -k "degrade or philox or b8" 2>&1 | tail -15
cat > /tmp/dg.py <<'PY'
import sys, torch; sys.path.insert(0, '.')
from csbsr_b200.data import degrade as G
from csbsr_b200.utils import synth
hr, _ = synth.batch(0, 8, 448); hr = hr.repeat(8,1,1,1).cuda()
p = torch.as_tensor(synth.degradation_params(64)).cuda()
for _ in range(3): G.degrade(hr, p)
torch.cuda.synchronize()
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:degrade --csv --log-file gpurun_out/degrade_launches.csv python /tmp/dg.py > /dev/null 2>&1
grep -o '"[a-z_:]*degrade[a-z_]*kernel[^"]*","[^"]*","[^"]*","[^"]*","gpu__time_duration.sum","[a-z]*","[0-9.,]*"' gpurun_out/degrade_launches.csv | tail -6
tail -4 gpurun_out/degrade_launches.csv
python scripts/bench_hbm_kernels.py 2>&1 | head -2
