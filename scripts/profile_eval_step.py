"""torch-profiler kernel table of one eval step (batch 64) -- live (not ncu-serialised) kernel times."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from csbsr_b200.config import cfg
from csbsr_b200.data import degrade as G
from csbsr_b200.engine import distributed as D, inference as E
from csbsr_b200.modeling.build_model import JointModel
from csbsr_b200.utils import synth
from torch.profiler import profile, ProfilerActivity
B = 64
c = cfg.clone(); c.merge_from_file("config/config_csbsr_pspnet.yaml")
m = JointModel(c); m.load_state_dict(synth.model_state_dict(), strict=True)
hr, mask = synth.batch(0, 16, 448)
hr = hr.repeat(4, 1, 1, 1).cuda(); mask = mask.repeat(4, 1, 1, 1).cuda()
params = torch.as_tensor(synth.degradation_params(B)).cuda()
def step():
    lr, _ = G.degrade(hr, params)
    sr, seg, kp = m(lr, None)
    for i in range(0, B, 16):
        E.seg_metrics(seg[i:i + 16], mask[i:i + 16], with_hd=True, to_host=False)
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=56))
