"""Achieved bandwidth of the HBM-bound kernels against their ALGORITHMIC bytes, CUDA-event timed,
inputs larger than L2 or rotated over several buffers.  Prints one line per kernel: time, GB/s, fraction of the measured peak."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from csbsr_b200 import kernels as K, _lib                      # noqa: E402
from csbsr_b200.data import degrade as G                       # noqa: E402
from csbsr_b200.engine import inference as E, losses as LS     # noqa: E402
from csbsr_b200.utils import synth                             # noqa: E402

PEAK = 6554.9
try:
    PEAK = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbps", PEAK)
except Exception:
    pass


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, ms, nbytes, note=""):
    gbs = nbytes / ms / 1e6
    print("%-34s %8.3f ms  %8.1f GB/s  %5.1f %% of %.0f  %s" % (name, ms, gbs, 100 * gbs / PEAK, PEAK, note))


B = 64
hr, mask = synth.batch(0, 8, 448)
hr = hr.repeat(8, 1, 1, 1).cuda(); mask = mask.repeat(8, 1, 1, 1).cuda()
params = torch.as_tensor(synth.degradation_params(B)).cuda()
ms = timed(lambda: G.degrade(hr, params))
report("csbsr_degrade (synth+blur+resize)", ms, B * 2560740, "FP32-FMA-bound: 531 MFLOP/img direct 21x21 blur = %.1f TFLOP/s" % (B * 0.531e9 / (ms * 1e-3) / 1e12))
sr = torch.rand(B, 3, 448, 448, device="cuda")
mean = torch.empty(B * 3, device="cuda"); rstd = torch.empty(B * 3, device="cuda")
ms = timed(lambda: K.clip_instnorm_stats(sr, mean, rstd, do_clip=True))
report("csbsr_clip_instnorm_stats", ms, B * 2 * 2408448, "read + clamped write + stats")
x = K.Fmap.empty(8, 448, 448, 128); x.t.normal_()
y = K.Fmap.empty(8, 448, 448, 128)
slope = torch.tensor([0.2], device="cuda")
ms = timed(lambda: _lib.lib().csbsr_prelu_fwd(x.ptr(), y.ptr(), slope.data_ptr(), x.t.numel(), _lib.stream_ptr()))
report("csbsr_prelu_fwd (8x448^2x128 bf16)", ms, 2 * x.t.numel() * 2)
gv = torch.empty(8, 128, device="cuda")
ms = timed(lambda: K.gap(x, gv, 128))
report("csbsr_gap_nhwc", ms, x.t.numel() * 2)
n = 89_100_000 // 4 * 4
p, g_, m, v = (torch.zeros(n, device="cuda") for _ in range(4))
ms = timed(lambda: _lib.lib().csbsr_adam_step(p.data_ptr(), g_.data_ptr(), m.data_ptr(), v.data_ptr(), n, 2e-5, 0.9, 0.999, 1e-8, 3, 1.0, 1, _lib.stream_ptr()))
report("csbsr_adam_step (89.1 M params)", ms, n * 32)
prob = torch.rand(16, 1, 448, 448, device="cuda")
m16 = mask[:16]
ms = timed(lambda: E.seg_metrics(prob, m16, with_hd=False, to_host=False), reps=3)
report("csbsr_seg_metrics AIU only (16 img)", ms, 16 * 1605632)
seg = torch.sigmoid(torch.randn(16, 1, 448, 448, device="cuda") * 3 - 2) * m16.clamp(0.05, 1)
ms = timed(lambda: E.seg_metrics(seg, m16, with_hd=True, to_host=False), reps=3)
report("csbsr_seg_metrics AIU+HD sweep", ms, 16 * 1605632, "%.0f (image x threshold) EDT+HD per s" % (16 * 99 / ms * 1e3))
ms = timed(lambda: LS.compute_sdf(m16), reps=3)
report("csbsr_sdf (16 masks)", ms, 16 * 2 * 802816)
