"""Timing of csbsr_blur_per_sample (21x21, stride 4) on 16 x 3 x 448^2 planes -- the KBlock pseudo-LR blur of the eval step."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from csbsr_b200 import kernels as K
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(16, 3, 448, 448, device="cuda", generator=g)
k = torch.rand(16, 441, device="cuda", generator=g); k = k / k.sum(1, keepdim=True)
lr = torch.rand(16, 3, 112, 112, device="cuda", generator=g)
err = torch.empty_like(lr)
for _ in range(3): K.blur_per_sample(x, k, lr, err, 21, 4)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): K.blur_per_sample(x, k, lr, err, 21, 4)
e1.record(); torch.cuda.synchronize()
print("blur_ps<21,4> 16 img: %.1f us" % (e0.elapsed_time(e1) * 100))
