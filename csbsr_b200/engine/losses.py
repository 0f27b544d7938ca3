"""Joint-training losses on the device (forward values + the gradient w.r.t. the segmentation predictions), with the
reference's semantics: BoundaryComboLoss / BoundaryLoss SDF / w^F / KBPNLoss / calc_loss
(model/utils/loss_functions.py, boundary_loss.py, oriented_weight.py, sr_loss_functions.py, model/engine/trainer.py:406-438).
All math runs in csrc/losses.cu (+ the blur / resize kernels for the pseudo-LR image); torch only holds the buffers."""
import ctypes as C

import torch

from .. import _lib
from ..data import degrade as G


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


def _f(t):
    return t.to(device=_dev(), dtype=torch.float32).contiguous()


def _ws(nbytes):
    return torch.empty(max(int(nbytes), 8), dtype=torch.uint8, device=_dev())


def compute_sdf(mask):
    """compute_sdf1_1 (boundary_loss.py:40-67) for a (B,1,H,W) mask -> fp32 (B,1,H,W) on the device."""
    m = _f(mask)
    b, c, h, w = m.shape
    assert c == 1
    out = torch.empty_like(m)
    L = _lib.lib()
    n = L.csbsr_sdf_workspace_bytes(b, h, w)
    ws = _ws(n)
    _lib.check(L.csbsr_sdf(m.data_ptr(), out.data_ptr(), b, h, w, ws.data_ptr(), n, _lib.stream_ptr()), "csbsr_sdf")
    _lib.count_launch("csbsr_sdf")
    return out


def seg_loss(p_main, p_aux, target, alpha, main_w=1.0, aux_w=0.4, sdf=None, upstream=None, need_grad=False):
    """out_map=False path: per-sample loss (B,) = main_w*L(main) + aux_w*L(aux); optionally d(sum upstream*loss)/dp."""
    pm, g = _f(p_main), _f(target)
    pa = _f(p_aux) if p_aux is not None else None
    b = pm.shape[0]
    hw = pm[0].numel()
    sd = sdf if sdf is not None else compute_sdf(g)
    loss = torch.empty(b, dtype=torch.float32, device=pm.device)
    gm = torch.empty_like(pm) if need_grad else None
    ga = torch.empty_like(pa) if (need_grad and pa is not None) else None
    up = _f(upstream) if upstream is not None else None
    L = _lib.lib()
    n = L.csbsr_seg_loss_workspace_bytes(b)
    ws = _ws(n)
    rc = L.csbsr_seg_loss(pm.data_ptr(), pa.data_ptr() if pa is not None else None, g.data_ptr(), sd.data_ptr(), b, hw,
                          C.c_float(alpha), C.c_float(main_w), C.c_float(aux_w), loss.data_ptr(),
                          gm.data_ptr() if gm is not None else None, ga.data_ptr() if ga is not None else None,
                          up.data_ptr() if up is not None else None, ws.data_ptr(), n, _lib.stream_ptr())
    _lib.check(rc, "csbsr_seg_loss")
    _lib.count_launch("csbsr_seg_loss")
    return (loss, gm, ga) if need_grad else loss


def seg_loss_wf_mean(p_main, p_aux, target, alpha, wf_amp, main_w=1.0, aux_w=0.4, sdf=None):
    """`.mean()` of the (B,B,H,W) loss tensor the reference forms when w^F is on: device fp64 scalar."""
    pm, g = _f(p_main), _f(target)
    pa = _f(p_aux) if p_aux is not None else None
    b, hw = pm.shape[0], pm[0].numel()
    sd = sdf if sdf is not None else compute_sdf(g)
    out = torch.empty(1, dtype=torch.float64, device=pm.device)
    L = _lib.lib()
    n = L.csbsr_seg_loss_wf_workspace_bytes(b, hw)
    ws = _ws(n)
    rc = L.csbsr_seg_loss_wf_mean(pm.data_ptr(), pa.data_ptr() if pa is not None else None, g.data_ptr(), sd.data_ptr(), b,
                                  hw, C.c_float(alpha), C.c_float(main_w), C.c_float(aux_w), C.c_float(wf_amp),
                                  out.data_ptr(), ws.data_ptr(), n, _lib.stream_ptr())
    _lib.check(rc, "csbsr_seg_loss_wf_mean")
    _lib.count_launch("csbsr_seg_loss_wf_mean")
    return out


def kbpn_loss(sr, hr, lr, kvec, k_gt, weights=(0.4, 0.4, 0, 2), ksize=21, factor=4):
    """KBPNLoss.forward (sr_loss_functions.py:39-56): kvec (B, ksize^2) is the (spatially constant) predicted kernel map.
    Returns (loss (B,), normalised kernel (B,1,k,k))."""
    from .. import kernels as K
    s, h, l = _f(sr), _f(hr), _f(lr)
    kv, kg = _f(kvec).view(s.shape[0], -1), _f(k_gt)
    kn = torch.empty_like(kv)
    K.vec_normalize(kv, kn)                                            # GAP of a constant map, then / sum (:85-87)
    blurred = torch.empty_like(s)
    K.blur_per_sample(s, kn, None, blurred, ksize, 1)                  # depthwise blur, stride 1 (:93)
    plr = G.FactorResize(factor, "bicubic")(blurred)                   # sr_transforms = antialiased bicubic (:94)
    b = s.shape[0]
    loss = torch.empty(b, dtype=torch.float32, device=s.device)
    L = _lib.lib()
    n = L.csbsr_sr_loss_workspace_bytes(b)
    ws = _ws(n)
    rc = L.csbsr_sr_loss(s.data_ptr(), h.data_ptr(), plr.data_ptr(), l.data_ptr(), kn.data_ptr(), kg.data_ptr(), b,
                         s[0].numel(), l[0].numel(), kn[0].numel(), C.c_float(weights[0]), C.c_float(weights[1]),
                         C.c_float(weights[2]), loss.data_ptr(), ws.data_ptr(), n, _lib.stream_ptr())
    _lib.check(rc, "csbsr_sr_loss")
    _lib.count_launch("csbsr_sr_loss")
    return loss, kn.view(b, 1, ksize, ksize)


def calc_loss(sr_loss, segment_loss_mean, task_loss_weight):
    """trainer.calc_loss (trainer.py:406-430): (1-beta)*mean(sr_loss) + beta*mean(segment_loss)."""
    return (1 - task_loss_weight) * sr_loss.mean() + task_loss_weight * segment_loss_mean
