"""Timing of the fused kernel-predictor chains vs the layer-by-layer path on 8 x 448^2 (one KBPN chunk)."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_model_gpu import _model_and_sd
from csbsr_b200 import kernels as K

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m, sd = _model_and_sd()
eng, _ = m._ensure_engines(torch.device("cuda", 0))
H = W = 448
sr_t = torch.rand(B, 3, H, W, device="cuda")
kvec = torch.rand(B, 441, device="cuda"); kvec /= kvec.sum(1, keepdim=True)
st = eng.p[0]


def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for fused in (True, False):
    eng.fused_kpred = fused
    ms = timeit(lambda: eng._kernel_predictor(st, sr_t, kvec, B, H, W, 0))
    print("kernel predictor (one stage, %d x 448^2) %s: %.3f ms" % (B, "fused chains" if fused else "layer-wise  ", ms))
a = eng.ws.fmap("hr_a64", B, H, W, 64)
ws = eng.ws.f32("kpred_ws", K._lib.lib().csbsr_kpred_workspace_bytes(B, H, W) // 4)
cb = torch.zeros(B, 5, 5, 64, device="cuda")
v = torch.zeros(B, 49, device="cuda")
t1 = timeit(lambda: K.kpred_sr_chain(sr_t, st["chain_sr"], a))
t2 = timeit(lambda: K.kpred_cat_chain(a, st["chain_cat"], cb, v, ws))
px = B * H * W
f1 = 2.0 * px * (27 * 49 + 49 * 32 + 2 * 288 * 32 + 288 * 49)
f2 = 2.0 * px * (49 * 32 + 288 * 32 + 288 * 49)
print("sr chain  %.3f ms  %.1f TFLOP/s useful" % (t1, f1 / t1 / 1e9))
print("cat chain %.3f ms  %.1f TFLOP/s useful" % (t2, f2 / t2 / 1e9))
