"""Debug: training graph vs fp32 oracle, stage by stage (GPU)."""
import os, sys
import numpy as np
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from csbsr_b200.modeling import params as P, train_graph as TG
from oracle import torch_ref as T

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
g = np.load("tests/golden/train_step.npz")
sd = P.synth_state_dict(P.kbpn_param_shapes(), prefix="sr_model.")
sd.update(P.synth_state_dict(P.pspnet_param_shapes(), prefix="segmentation_model."))
sd = {k: v.cuda() for k, v in sd.items()}
lr = torch.from_numpy(g["lr"]).cuda()
torch.backends.cudnn.enabled = False
with torch.no_grad():
    sr, kvec = TG.kbpn_forward(sd, lr)
    T.BN_TRAIN = True
    sr_ref, kvec_ref = T.kbpn_forward(sd, lr)
    print("sr diff", (sr - sr_ref).abs().max().item(), "kvec", (kvec - kvec_ref.view(2, -1)).abs().max().item(), kvec_ref.abs().max().item())
    x = F.instance_norm(sr_ref, eps=1e-5)
    for train in (False, True):
        T.BN_TRAIN = train
        sd2 = {k: v.clone() for k, v in sd.items()}
        seg, aux = TG.pspnet_forward(sd2, x, bn_training=train, dropout=False)
        seg_ref, aux_ref = T.pspnet_forward(sd, x)
        print("bn_train", train, "seg diff max", (seg - seg_ref).abs().max().item(), "mean", (seg - seg_ref).abs().mean().item(),
              "aux", (aux - aux_ref).abs().max().item(), (aux - aux_ref).abs().mean().item())

# per-BN comparison in train mode
rec_a, rec_b = [], []
_tg_bn, _t_bn = TG._bn, T._bn
def tg_bn(P_, p, x, training, momentum=0.1):
    y = _tg_bn(P_, p, x, training, momentum); rec_a.append((p, x.detach(), y.detach())); return y
def t_bn(sd_, p, x):
    y = _t_bn(sd_, p, x); rec_b.append((p, x.detach(), y.detach())); return y
TG._bn, T._bn = tg_bn, t_bn
T.BN_TRAIN = True
with torch.no_grad():
    sd2 = {k: v.clone() for k, v in sd.items()}
    TG.pspnet_forward(sd2, x, bn_training=True, dropout=False)
    T.pspnet_forward(sd, x)
for (pa, xa, ya), (pb, xb, yb) in zip(rec_a, rec_b):
    c = xb.shape[1]
    xa_, ya_ = xa[..., :c].permute(0, 3, 1, 2).float(), ya[..., :c].permute(0, 3, 1, 2).float()
    print("%-50s in rel %.4f out rel %.4f  | min var %.3e" % (pa[19:], (xa_ - xb).abs().max().item() / xb.abs().max().item(),
          (ya_ - yb).abs().max().item() / yb.abs().max().item(), xb.var(dim=(0, 2, 3), unbiased=False).min().item()))
