"""TEST INFRASTRUCTURE ONLY -- torch/numpy restatement of the reference's joint-training losses (forward), pinned
against the UNMODIFIED reference classes by tests/golden/losses.npz (tests/golden/gen_golden.py).

  sdf                 compute_sdf1_1                       model/utils/boundary_loss.py:40-67
  boundary_combo      BoundaryComboLoss.forward            model/utils/loss_functions.py:49-74 (+ :196-210, :284-345)
  wf_weight           SegmentFailerOrientedExpWeight       model/utils/oriented_weight.py:73-83
  seg_loss            MetaSSLossCalc.calc_ss_loss + JointModelWithLoss.multiple_weight   model/modeling/build_model.py:258-278,422-438
  kbpn_loss           KBPNLoss.forward + Get_pseudo_lr     model/utils/sr_loss_functions.py:39-56,84-102

`find_boundaries(mode='inner')` comes from scikit-image, which is absent from this image and un-pinned in the
reference's requirement.txt: its 4-connected inner-boundary definition is restated here (parity unpinned at this
one boundary -- it only zeroes the SDF on the mask's inner contour)."""
import numpy as np
import torch
import torch.nn.functional as F
from scipy import ndimage


def sdf(mask):
    """mask (B,1,H,W) float -> normalised signed distance map (B,1,H,W) float32."""
    img = np.asarray(mask).astype(np.uint8)
    out = np.zeros(img.shape)
    fp = ndimage.generate_binary_structure(3, 1)
    for b in range(img.shape[0]):
        pos = img[b].astype(bool)                       # (1,H,W): the reference runs the EDT on this 3-D array
        if pos.any():
            neg = ~pos
            posdis = ndimage.distance_transform_edt(pos)
            negdis = ndimage.distance_transform_edt(neg)
            u8 = pos.astype(np.uint8)
            boundary = (ndimage.grey_dilation(u8, footprint=fp) != ndimage.grey_erosion(u8, footprint=fp)) & (u8 != 0)
            s = (negdis - np.min(negdis)) / (np.max(negdis) - np.min(negdis)) - \
                (posdis - np.min(posdis)) / (np.max(posdis) - np.min(posdis))
            s[boundary] = 0
            out[b][0] = s[0]
    return torch.from_numpy(out).float()


def _wbce(p, g, reduce):
    loss = -(1 * g * torch.log(p + 1e-8) + 1 * (1 - g) * torch.log(1 - p + 1e-8)) / 2
    return loss.mean(dim=(1, 2, 3)) if reduce else loss


def _dice(p, g, out_map):
    if out_map:
        num = 2 * torch.sum(p * g, dim=1) + 1e-6
        den = torch.sum(p.pow(2) + g.pow(2)) + 1e-6
        return 1 / g.numel() - num / den
    pf, gf = p.contiguous().view(p.shape[0], -1), g.contiguous().view(g.shape[0], -1)
    return 1 - (2 * torch.sum(pf * gf, dim=1) + 1e-6) / (torch.sum(pf.pow(2) + gf.pow(2), dim=1) + 1e-6)


def boundary_combo(p, g, sdf_map, alpha, out_map=False):
    p = p.clamp(min=1e-8)
    wd = (_wbce(p, g, not out_map) + _dice(p, g, out_map)) / 2
    bd = p * sdf_map
    if not out_map:
        bd = bd.mean(dim=(1, 2, 3))
    return alpha * wd + (1 - alpha) * bd


def wf_weight(p, g, amp):
    return torch.exp(amp * torch.abs(p.detach() - g))


def seg_loss(p_main, p_aux, g, alpha, main_w=1.0, aux_w=0.4, wf_amp=0.0):
    """-> per-sample loss (B,) when wf_amp == 0, else the (B,B,H,W) tensor of the reference."""
    s = sdf(g.detach().cpu().numpy()).to(p_main.device)
    out_map = wf_amp != 0
    loss = main_w * boundary_combo(p_main, g, s, alpha, out_map) + aux_w * boundary_combo(p_aux, g, s, alpha, out_map)
    if out_map:
        loss = wf_weight(p_main, g, wf_amp) * loss
    return loss


def kbpn_loss(sr, hr, lr, kernel_map, k_gt, weights=(0.4, 0.4, 0, 2), ksize=21, factor=4):
    """-> (loss (B,), pseudo_lr, normalised kernel (B,1,k,k))."""
    k = kernel_map.mean(dim=(2, 3), keepdim=True)
    k = k / k.sum(dim=1).view(-1, 1, 1, 1)
    w = k.view(-1, 1, ksize, ksize)
    outs = []
    for i in range(sr.shape[0]):
        t = F.conv2d(sr[i:i + 1], w[i].expand(3, 1, ksize, ksize), padding=(ksize - 1) // 2, groups=3)
        outs.append(F.interpolate(t, size=(sr.shape[2] // factor, sr.shape[3] // factor), mode="bicubic", antialias=True,
                                  align_corners=False))
    plr = torch.cat(outs, 0)
    loss = weights[0] * (sr - hr).abs().mean((1, 2, 3)) + weights[1] * (plr - lr).abs().mean((1, 2, 3)) + \
        weights[2] * ((w - k_gt) ** 2).mean((1, 2, 3))
    return loss, plr, w
