"""TEST INFRASTRUCTURE ONLY -- fp32 PyTorch restatement of one CSBSR training step's forward + loss
(JointModelWithLoss.forward at an iteration where every phase is active, model/modeling/build_model.py:390-416;
calc_loss, model/engine/trainer.py:406-438).  Gradients come from torch autograd on this graph.

Differences from the eval oracle (oracle/torch_ref.py): BatchNorm uses batch statistics, the SR output is NOT
clipped before the instance norm (clip_sr is eval-only, build_model.py:143-146), Dropout2d is disabled (the parity
harness of SURVEY section 7: masks are random in the reference).  Pinned against the unmodified reference by
tests/golden/train_step.npz (tests/golden/gen_golden.py train).
"""
import torch
import torch.nn.functional as F

from . import loss_ref as L
from . import torch_ref as T


def train_forward(sd, lr, hr, mask, kgt, alpha, beta=0.3, wf_amp=1.0, bn_train=True, hrnet=False, gt_kernel_phase=False,
                  sr_only=False, blur_skip=False):
    """-> (loss, seg_loss, sr_loss, sr, seg, aux).  `sd` tensors that require grad receive gradients."""
    T.BN_TRAIN = bn_train
    try:
        sr, kvec = T.kbpn_forward(sd, lr, gt_kernel=kgt if gt_kernel_phase else None)
        if blur_skip:
            seg, aux = T.pspnet_forward(sd, F.instance_norm(sr, eps=1e-5), kvec=kvec)
        else:
            seg, aux = (T.hrnet_ocr_forward if hrnet else T.pspnet_forward)(sd, F.instance_norm(sr, eps=1e-5))
    finally:
        T.BN_TRAIN = False
    kmap = kvec.expand(-1, -1, lr.shape[2], lr.shape[3])
    sr_loss, _, _ = L.kbpn_loss(sr, hr, lr, kmap, kgt)
    seg_loss = L.seg_loss(seg, aux, mask, alpha, wf_amp=wf_amp)
    loss = sr_loss.mean() if sr_only else (1 - beta) * sr_loss.mean() + beta * seg_loss.mean()   # calc_pretrain_loss
    return loss, seg_loss, sr_loss, sr, seg, aux
