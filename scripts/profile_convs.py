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
for label, flops, e0, e1, _u in K.PROFILE:
    t = e0.elapsed_time(e1)
    a = agg.setdefault(label, [0, 0.0, 0.0])
    a[0] += 1; a[1] += t; a[2] += flops
tot = sum(a[1] for a in agg.values())
print("total conv ms", tot, "per img", tot / B)
for label, (n, t, f) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-50s x%3d %8.3f ms %5.1f%%  %7.1f TFLOP/s(padded)" % (label, n, t, 100 * t / tot, f / t / 1e9))
