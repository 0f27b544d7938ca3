import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CSBSR_CONV_TRACE"] = "1"
import ctypes, numpy as np, torch
from csbsr_b200 import kernels as K, _lib
which = sys.argv[1] if len(sys.argv) > 1 else "hr1"
B = 8
g = torch.Generator(device="cuda").manual_seed(0)
if which == "hr1":
    x = K.Fmap.empty(B, 448, 448, 64); x.t.normal_()
    pc = K.pack_conv(torch.randn(64, 64, 1, 1, device="cuda") * 0.1); y = K.Fmap.empty(B, 448, 448, 64); kw = dict(act=K.ACT_LEAKY, slope=0.01)
elif which == "ikc3":
    x = K.Fmap.empty(B, 448, 448, 64); x.t.normal_()
    pc = K.pack_conv(torch.randn(64, 64, 3, 3, device="cuda") * 0.05, padding=1); y = K.Fmap.empty(B, 448, 448, 64); kw = dict(act=K.ACT_LEAKY, slope=0.01)
elif which == "ikc32":
    x = K.Fmap.empty(B, 448, 448, 64); x.t.normal_()
    pc = K.pack_conv(torch.randn(32, 32, 3, 3, device="cuda") * 0.05, padding=1, cout_pad=64, cin_pad=32); y = K.Fmap.empty(B, 448, 448, 64); kw = dict(act=K.ACT_LEAKY, slope=0.01)
elif which == "n16":
    x = K.Fmap.empty(B, 448, 448, 128); x.t.normal_()
    pc = K.pack_conv(torch.randn(12, 128, 3, 3, device="cuda") * 0.05, padding=1); y = torch.zeros(B, 12, 448, 448, device="cuda"); kw = dict()
elif which == "c8s4":
    x = K.Fmap.empty(B, 448, 448, 128); x.t.normal_()
    pc = K.pack_conv(torch.randn(128, 128, 8, 8, device="cuda") * 0.01, stride=4, padding=2)
    r = K.Fmap.empty(B, 112, 112, 128); r.t.normal_(); y = K.Fmap.empty(B, 112, 112, 128)
    kw = dict(act=K.ACT_LEAKY, slope=0.1, r1=r, r1_sign=-1.0)
elif which == "sft1":
    x = K.Fmap.empty(B, 112, 112, 832); x.t.normal_()
    pc = K.pack_conv(torch.randn(384, 832, 3, 3, device="cuda") * 0.01, padding=1); y = K.Fmap.empty(B, 112, 112, 384); kw = dict()
else:
    x = K.Fmap.empty(B, 112, 112, 128); x.t.normal_()
    pc = K.pack_deconv8s4(torch.randn(128, 128, 8, 8, device="cuda") * 0.02); y = K.Fmap.empty(B, 448, 448, 128)
    r = K.Fmap.empty(B, 448, 448, 128); r.t.normal_(); kw = dict(act=K.ACT_LEAKY, slope=0.1, r1=r)
for _ in range(3): K.conv(x, pc, y, **kw)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (10 * 256))()
L = ctypes.CDLL(_lib.LIB_PATH)
assert L.csbsr_conv_trace_read(buf) == 0
t = np.array(buf[:]).reshape(10, 256).astype(np.int64)
n = int((t[5] > 0).sum())
t0 = t[0, 0]
names = ["prod_issue", "mma_acc_free", "mma_first", "mma_commit", "epi_start", "epi_store"]
print("tiles on CTA0:", n)
for i in list(range(0, 6)) + list(range(20, 26)):
    print(i, " ".join("%s=%d" % (names[k], t[k, i] - t0) for k in range(6)))
d = np.diff(t[5, 10:n - 2]); print("steady per-tile cycles (store to store): mean %.0f min %d max %d" % (d.mean(), d.min(), d.max()))
print("MMA warp: cycles waiting on full barriers per tile: mean %.0f ; tile span acc_free->commit mean %.0f" % (t[2, 10:n-2].mean(), (t[3, 10:n-2] - t[1, 10:n-2]).mean()))
print("epi duration mean", (t[5, 10:n-2] - t[4, 10:n-2]).mean(), " mma commit->epi start", (t[4, 10:n-2] - t[3, 10:n-2]).mean(),
      " acc_free->first mma", (t[2, 10:n-2] - t[1, 10:n-2]).mean(), " first mma->commit", (t[3, 10:n-2] - t[2, 10:n-2]).mean(),
      " prod lead over mma_first", (t[2, 10:n-2] - t[0, 10:n-2]).mean())

# staged epilogue breakdown of the leader warp (rows 6..9 need the -DCSBSR_CONV_TRACE_BUILD build)
if (t[6, 10:n - 2] > 0).all():
    sl = slice(10, n - 2)
    print("epilogue leader: loop top -> store-read wait + team barrier %.0f | -> accumulator ready %.0f | math %.0f | fence.proxy %.0f | "
          "team barrier -> store issued %.0f | store issued -> next loop top (other work / idle) %.0f" % (
              (t[7, sl] - t[6, sl]).mean(), (t[4, sl] - t[7, sl]).mean(), (t[8, sl] - t[4, sl]).mean(), (t[9, sl] - t[8, sl]).mean(),
              (t[5, sl] - t[9, sl]).mean(), (t[6, 12:n - 2] - t[5, 10:n - 4]).mean()))
