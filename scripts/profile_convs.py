import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import collections
import torch
from tests.test_model_gpu import _model_and_sd
from csbsr_b200 import kernels as K
m, sd = _model_and_sd()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m.chunk = B
x = torch.rand(B, 3, 112, 112).cuda()
dk = torch.zeros(B, 1, 7, 7)
for _ in range(2):
    m(x, dk)
torch.cuda.synchronize()
K.PROFILE = []
m(x, dk)
torch.cuda.synchronize()
agg = collections.OrderedDict()
for label, flops, e0, e1, _u, nbytes in K.PROFILE:
    t = e0.elapsed_time(e1)
    a = agg.setdefault(label, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += t; a[2] += flops; a[3] += nbytes
tot = sum(a[1] for a in agg.values())
print("total conv ms", tot, "per img", tot / B)
ideal_tot = 0.0
for label, (n, t, f, nb) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    t_tc, t_hbm = f / 1400.5e9, nb / 6554.9e6          # ms at the measured peaks
    ideal = max(t_tc, t_hbm)
    ideal_tot += ideal
    print("%-48s x%3d %7.3f ms %5.1f%% %7.1f TF/s %6.0f GB/s  ideal %6.3f (%s) excess %6.3f" %
          (label, n, t, 100 * t / tot, f / t / 1e9, nb / t / 1e6, ideal, "tc" if t_tc > t_hbm else "hbm", t - ideal))
print("sum of per-layer roofline times %.3f ms (%.0f%% of measured)" % (ideal_tot, 100 * ideal_tot / tot))
