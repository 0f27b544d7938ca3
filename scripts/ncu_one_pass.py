"""One profiled pass for ncu (`--profile-from-start off`): a warm pass, then cudaProfilerStart .. one pass .. cudaProfilerStop.

  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      --csv --log-file out.csv python scripts/ncu_one_pass.py eval|train [images]
eval : degrade -> KBPN -> PSPNet -> AIU + HD/MSD on `images` (default 16) 448^2 images, eager launches (chunk 16 / 32 / 16)
train: one joint training step (batch 8, 224^2 crops, iteration 40000) incl. the fused Adam launch"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from csbsr_b200.config import cfg
from csbsr_b200.data import degrade as G
from csbsr_b200.engine import inference as E
from csbsr_b200.utils import synth

what = sys.argv[1] if len(sys.argv) > 1 else "eval"
c = cfg.clone()
c.merge_from_file("config/config_csbsr_pspnet.yaml")
if what == "eval":
    from csbsr_b200.modeling.build_model import JointModel
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    m = JointModel(c)
    m.load_state_dict(synth.model_state_dict(), strict=True)
    hr, mask = synth.batch(0, B, 448)
    hr, mask = hr.cuda(), mask.cuda()
    params = torch.as_tensor(synth.degradation_params(B)).cuda()

    def run():
        lr, _ = G.degrade(hr, params)
        sr, seg, kp = m(lr, None)
        for i in range(0, B, 16):
            E.seg_metrics(seg[i:i + 16], mask[i:i + 16], with_hd=True, to_host=False)
else:
    from csbsr_b200.engine.optim import FusedAdam
    from csbsr_b200.engine.trainer import train_step
    from csbsr_b200.modeling.build_model import JointModelWithLoss
    c.SOLVER.SEG_FAIL_ORIENTED_WEIGHT4SS_AMP = 1.0
    m = JointModelWithLoss(c, num_train_ds=1000, resume_iter=40000)
    m.load_state_dict(synth.model_state_dict(), strict=True)
    m.cuda().train()
    opt = FusedAdam(m.parameters(), lr=c.SOLVER.LR)
    hr, mask = synth.batch(1000, 8, 224)
    hr, mask = hr.cuda(), mask.cuda()
    params = torch.as_tensor(synth.degradation_params(8, seed=50)).cuda()
    it = [40000]

    def run():
        it[0] += 1
        train_step(m, opt, c, it[0], hr, mask, params, 1)

run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one %s pass" % what)
