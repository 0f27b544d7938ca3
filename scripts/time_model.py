import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time
import torch
from tests.test_model_gpu import _model_and_sd
m, sd = _model_and_sd()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 8
m.chunk = chunk
x = torch.rand(B, 3, 112, 112).cuda()
dk = torch.zeros(B, 1, 7, 7)
for _ in range(2):
    out = m(x, dk)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
sr_eng, ss_eng = m._ensure_engines(torch.device("cuda", 0))
xc = x[:chunk].contiguous()
for name, fn in (("kbpn", lambda: sr_eng.forward(xc)), ("full", lambda: m(x, dk))):
    torch.cuda.synchronize(); e0.record()
    for _ in range(3): fn()
    e1.record(); torch.cuda.synchronize()
    n = chunk if name == "kbpn" else B
    ms = e0.elapsed_time(e1) / 3
    print(name, "ms/iter", ms, "ms/img", ms / n, "img/s", n / ms * 1e3)
sr = sr_eng.forward(xc)[0]
mean = torch.empty(chunk * 3, device="cuda"); rstd = torch.empty(chunk * 3, device="cuda")
torch.cuda.synchronize(); e0.record()
for _ in range(3): ss_eng.forward(sr, mean.fill_(0.5), rstd.fill_(2.0))
e1.record(); torch.cuda.synchronize()
print("pspnet ms/img", e0.elapsed_time(e1) / 3 / chunk)
print("mem GB", torch.cuda.max_memory_allocated() / 1e9)
