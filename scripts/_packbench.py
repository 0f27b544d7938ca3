import sys, torch; sys.path.insert(0, '.')
from csbsr_b200 import kernels as K
from csbsr_b200.engine.optim import FusedAdam
g = torch.Generator().manual_seed(0)
shapes = [(128, 128, 8, 8)] * 20 + [(569, 569, 3, 3), (697, 697, 3, 3), (825, 825, 3, 3)] * 2 + [(256, 256, 3, 3)] * 12 + [(512, 512, 3, 3)] * 5
ws = [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in shapes]
opt = FusedAdam(ws, lr=1e-3)
def pack_all():
    for w in ws:
        a, b, R, S = w.shape
        K._pack_device(w, (a + 63) // 64 * 64, (b + 63) // 64 * 64, 0)
        K._pack_device(w, (b + 63) // 64 * 64, (a + 63) // 64 * 64, 1 if R == 3 else 2)
pack_all()
n = sum(w.numel() for w in ws)
for it in range(3):
    for p in ws: p.grad.normal_()
    opt.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pack_all(); e1.record(); torch.cuda.synchronize()
    print("multi pack of %.1f M params x 2 layouts: %.3f ms (tiles %d)" % (n / 1e6, e0.elapsed_time(e1), K._PACK_STATE["total"]))
