"""Standalone timing of the conv kernel on the layer shapes that dominate the KBPN/PSPNet step (8 images)."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from csbsr_b200 import kernels as K

def run(name, x, pc, y, iters=5, **kw):
    for _ in range(2): K.conv(x, pc, y, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): K.conv(x, pc, y, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * x.n * (y.h if isinstance(y, K.Fmap) else y.shape[2]) * (y.w if isinstance(y, K.Fmap) else y.shape[3]) * pc.macs_per_pixel * (1 if pc.os == 1 else 1)
    print("%-28s %8.3f ms  %8.1f TFLOP/s useful" % (name, ms, flops / ms / 1e9))

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
only = sys.argv[2] if len(sys.argv) > 2 else ""
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
cases = []
if not only or only == "ikc3":
    x = K.Fmap.empty(B, 448, 448, 64); x.t.normal_()
    pc = K.pack_conv(rn(64, 64, 3, 3) * 0.05, padding=1)
    run("ikc 3x3 64->64 @448", x, pc, K.Fmap.empty(B, 448, 448, 64), act=K.ACT_LEAKY, slope=0.01)
if not only or only == "hr1":
    x = K.Fmap.empty(B, 448, 448, 64); x.t.normal_()
    pc = K.pack_conv(rn(64, 64, 1, 1) * 0.1)
    run("1x1 64->64 @448", x, pc, K.Fmap.empty(B, 448, 448, 64), act=K.ACT_LEAKY, slope=0.01)
if not only or only == "deconv":
    x = K.Fmap.empty(B, 112, 112, 128); x.t.normal_()
    pc = K.pack_deconv8s4(rn(128, 128, 8, 8) * 0.02)
    y = K.Fmap.empty(B, 448, 448, 128); r = K.Fmap.empty(B, 448, 448, 128); r.t.normal_()
    run("deconv 128->128 +res", x, pc, y, act=K.ACT_LEAKY, slope=0.1, r1=r)
    run("deconv 128->128", x, pc, y, act=K.ACT_LEAKY, slope=0.1)
if not only or only == "c8s4":
    x = K.Fmap.empty(B, 448, 448, 128); x.t.normal_()
    pc = K.pack_conv(rn(128, 128, 8, 8) * 0.01, stride=4, padding=2)
    r = K.Fmap.empty(B, 112, 112, 128); r.t.normal_()
    run("conv8s4 128->128 -res", x, pc, K.Fmap.empty(B, 112, 112, 128), act=K.ACT_LEAKY, slope=0.1, r1=r, r1_sign=-1.0)
if not only or only == "n16":
    x = K.Fmap.empty(B, 448, 448, 512); x.t.normal_()
    pc = K.pack_conv(rn(3, 512, 3, 3) * 0.01, padding=1)
    run("3x3 512->3 @448 (f32)", x, pc, torch.empty(B, 3, 448, 448, device="cuda"))
if not only or only == "sft":
    x = K.Fmap.empty(B, 112, 112, 384); x.t.normal_()
    pc = K.pack_conv(rn(825, 384, 3, 3) * 0.01, padding=1, cout_pad=832)
    run("sft0 3x3 384->832 @112", x, pc, K.Fmap.empty(B, 112, 112, 832), act=K.ACT_LEAKY, slope=0.1)
    x = K.Fmap.empty(B, 112, 112, 832); x.t.normal_()
    pc = K.pack_conv(rn(384, 832, 3, 3) * 0.01, padding=1)
    run("sft1 3x3 832->384 @112", x, pc, K.Fmap.empty(B, 112, 112, 384))
