"""HD / MSD sweep timing on saved network maps (build/bench_seg16.pt, written by a bench run) and on the App. E noisy-crack fixture."""
import os, sys, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "scripts"))
from csbsr_b200.engine import inference as E
if os.path.exists("build/bench_seg16.pt"):                 # maps saved from an earlier run (build/ is git-ignored)
    d = torch.load("build/bench_seg16.pt")
    seg, mask = d["seg"].cuda(), d["mask"].cuda()
else:                                                       # the bench's own maps: synthetic weights on 16 synthetic images
    from csbsr_b200.config import cfg
    from csbsr_b200.data import degrade as G
    from csbsr_b200.modeling.build_model import JointModel
    from csbsr_b200.utils import synth
    c = cfg.clone(); c.merge_from_file("config/config_csbsr_pspnet.yaml")
    m = JointModel(c); m.load_state_dict(synth.model_state_dict(), strict=True)
    hr, mask = synth.batch(0, 16, 448)
    lr, _ = G.degrade(hr.cuda(), torch.as_tensor(synth.degradation_params(16, seed=5)).cuda())
    with torch.no_grad():
        _, seg, _ = m(lr, None)
    mask = mask.cuda()
    os.makedirs("build", exist_ok=True)
    torch.save({"seg": seg.cpu(), "mask": mask.cpu()}, "build/bench_seg16.pt")
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("network maps, 16 img: %.3f ms" % t(lambda: E.seg_metrics(seg, mask, with_hd=True, to_host=False)))
import bench_hbm_kernels as B
pn, mn = B.noisy_crack_case(16)
pn, mn = pn.cuda(), mn.cuda()
print("App. E noisy crack, 16 img: %.3f ms" % t(lambda: E.seg_metrics(pn, mn, with_hd=True, to_host=False)))
