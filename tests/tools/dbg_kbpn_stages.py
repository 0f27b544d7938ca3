import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tests.test_model_gpu import _model_and_sd
from oracle import torch_ref as T
import torch.nn.functional as F
m, sd = _model_and_sd()
sdc = {k: v.cuda() for k, v in sd.items()}
g = torch.Generator().manual_seed(5)
x = torch.rand(2, 3, 24, 32, generator=g).cuda()
sr_eng, _ = m._ensure_engines(torch.device("cuda", 0))
sr_eng.debug = {}
sr, kvec = sr_eng.forward(x)
with torch.no_grad():
    sr_ref, kvec_ref, inter = T.kbpn_forward(sdc, x, return_intermediates=True)
    # oracle with bf16-rounded weights & activations rounding emulation is not available; just print raw sums
    for k in ["init_f","init_kernel","sr_t0","kvec0","sr_t1","kvec1","sr_t2","kvec2","sr_t3","kvec3"]:
        a = sr_eng.debug[k]; b = inter[k].reshape(a.shape)
        print(k, "max-abs", float((a-b).abs().max()), "ref max", float(b.abs().max()))
    # raw (unnormalised) sums in the oracle: recompute delta sums
    kv = inter["init_kernel"]
    for s in range(4):
        p = "sr_model.back_projection_stages.%d.kb.kernel_predictor" % s
        d = T.kernel_predictor_ikc(sdc, p, inter["sr_t%d"%s], kv, 21)
        print("stage", s, "raw sum ref", d.sum(dim=1).flatten().tolist())
        kv = inter["kvec%d"%s]
