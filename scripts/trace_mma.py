import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CSBSR_CONV_TRACE"] = "1"
import ctypes, numpy as np, torch
from csbsr_b200 import kernels as K, _lib
L = ctypes.CDLL(_lib.LIB_PATH)
def trace():
    buf = (ctypes.c_longlong * (6 * 256))()
    assert L.csbsr_conv_trace_read(buf) == 0
    return np.array(buf[:]).reshape(6, 256).astype(np.int64)
B = 4
for cin, cout, k, hw in [(64, 16, 3, 448), (64, 64, 3, 448), (64, 128, 3, 448), (64, 256, 3, 448), (128, 128, 1, 448), (256, 256, 3, 112), (64, 64, 1, 448)]:
    x = K.Fmap.empty(B, hw, hw, cin); x.t.normal_()
    pc = K.pack_conv(torch.randn(cout, cin, k, k, device="cuda") * 0.05, padding=k // 2)
    y = K.Fmap.empty(B, hw, hw, max(cout, 64)) if cout >= 64 else K.Fmap.empty(B, hw, hw, 16)
    for _ in range(2): K.conv(x, pc, y)
    torch.cuda.synchronize()
    t = trace(); n = int((t[3] > 0).sum())
    nm = k * k * (cin // 64) * 4
    span = (t[3, 5:n-2] - t[1, 5:n-2]).mean(); wait = t[2, 5:n-2].mean()
    d = np.diff(t[3, 5:n-2]).mean()
    print("cin %d cout %d k%d: MMAs/tile %d, issue span %.0f, barrier wait %.0f -> %.1f cycles/MMA ; tile period %.0f" % (cin, cout, k, nm, span, wait, (span - wait) / nm, d))
