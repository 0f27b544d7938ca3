"""Which conv layers are power-limited?  Loops one layer shape for ~1.5 s while sampling SM clock and board power
(pynvml), then prints the achieved TFLOP/s next to the median clock / power under load."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pynvml
import torch
from csbsr_b200 import kernels as K

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.stop, self.clk, self.pw = False, [], []

    def run(self):
        while not self.stop:
            self.clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
            time.sleep(0.02)


def probe(name, x, pc, y, secs=1.5, **kw):
    for _ in range(3):
        K.conv(x, pc, y, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); K.conv(x, pc, y, **kw); e1.record(); torch.cuda.synchronize()
    iters = max(10, int(secs * 1000 / max(e0.elapsed_time(e1), 1e-3)))
    s = Sampler(); s.start()
    e0.record()
    for _ in range(iters):
        K.conv(x, pc, y, **kw)
    e1.record(); torch.cuda.synchronize()
    s.stop = True; s.join()
    ms = e0.elapsed_time(e1) / iters
    yh = y.h if isinstance(y, K.Fmap) else y.shape[2]
    yw = y.w if isinstance(y, K.Fmap) else y.shape[3]
    fl = 2.0 * x.n * yh * yw * pc.macs_per_pixel
    k = len(s.clk) // 3
    print("%-26s %7.3f ms %7.1f TF/s  first-launch %7.3f ms | clk MHz median %4.0f min %4.0f | power W median %4.0f max %4.0f"
          % (name, ms, fl / ms / 1e9, e0.elapsed_time(e1) / iters, np.median(s.clk[k:]), min(s.clk[k:]), np.median(s.pw[k:]), max(s.pw)))


B = 8
rn = lambda *s: torch.randn(*s, device="cuda")
x = K.Fmap.empty(B, 448, 448, 128); x.t.normal_()
pc = K.pack_conv(rn(128, 128, 8, 8) * 0.01, stride=4, padding=2)
probe("conv8s4 128->128", x, pc, K.Fmap.empty(B, 112, 112, 128), act=K.ACT_LEAKY, slope=0.1)
x = K.Fmap.empty(B, 112, 112, 128); x.t.normal_()
pc = K.pack_deconv8s4(rn(128, 128, 8, 8) * 0.02)
probe("deconv 128->128", x, pc, K.Fmap.empty(B, 448, 448, 128), act=K.ACT_LEAKY, slope=0.1)
x = K.Fmap.empty(B, 112, 112, 832); x.t.normal_()
pc = K.pack_conv(rn(384, 832, 3, 3) * 0.01, padding=1)
probe("sft1 3x3 832->384", x, pc, K.Fmap.empty(B, 112, 112, 384))
x = K.Fmap.empty(B, 448, 448, 64); x.t.normal_()
pc = K.pack_conv(rn(32, 32, 3, 3) * 0.05, padding=1, cout_pad=64, cin_pad=32)
probe("3x3 32->32 @448", x, pc, K.Fmap.empty(B, 448, 448, 64), act=K.ACT_LEAKY, slope=0.01)
pc = K.pack_conv(rn(64, 64, 1, 1) * 0.1)
probe("1x1 64->64 @448", x, pc, K.Fmap.empty(B, 448, 448, 64), act=K.ACT_LEAKY, slope=0.01)
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16); b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(3): a @ b
torch.cuda.synchronize()
s = Sampler(); s.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(1500): a @ b
e1.record(); torch.cuda.synchronize(); s.stop = True; s.join()
ms = e0.elapsed_time(e1) / 1500; k = len(s.clk) // 3
print("cuBLAS bf16 8192^3         %7.3f ms %7.1f TF/s | clk MHz median %4.0f min %4.0f | power W median %4.0f max %4.0f" % (
    ms, 2 * 8192 ** 3 / ms / 1e9, np.median(s.clk[k:]), min(s.clk[k:]), np.median(s.pw[k:]), max(s.pw)))
