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


def calc_loss(sr_loss, segment_loss_mean, task_loss_weight, iteration=None, cfg=None):
    """trainer.calc_loss + calc_pretrain_loss (trainer.py:406-438): (1-beta)*mean(sr_loss) + beta*mean(segment_loss); only
    the SR term while `iteration` is inside SOLVER.SR_PRETRAIN_ITER, only the segmentation term inside SEG_PRETRAIN_ITER."""
    loss = (1 - task_loss_weight) * sr_loss.mean() + task_loss_weight * segment_loss_mean
    if iteration is not None and cfg is not None:
        if cfg.SOLVER.SR_PRETRAIN_ITER[0] <= iteration < cfg.SOLVER.SR_PRETRAIN_ITER[1]:
            loss = sr_loss.mean()
        if cfg.SOLVER.SEG_PRETRAIN_ITER[0] <= iteration < cfg.SOLVER.SEG_PRETRAIN_ITER[1]:
            loss = segment_loss_mean
    return loss


# ------------------------------------------------------------------ differentiable forms used by the training step
class _SegLossFn(torch.autograd.Function):
    """Per-sample BoundaryCombo loss (out_map=False) on the fused kernel; backward re-runs it with the upstream weights."""

    @staticmethod
    def forward(ctx, p_main, p_aux, target, sdf, alpha, main_w, aux_w):
        ctx.save_for_backward(p_main, p_aux, target, sdf)
        ctx.cfg = (alpha, main_w, aux_w)
        return seg_loss(p_main, p_aux, target, alpha, main_w, aux_w, sdf=sdf)

    @staticmethod
    def backward(ctx, up):
        p_main, p_aux, target, sdf = ctx.saved_tensors
        alpha, main_w, aux_w = ctx.cfg
        _, gm, ga = seg_loss(p_main, p_aux, target, alpha, main_w, aux_w, sdf=sdf, upstream=up, need_grad=True)
        return gm, ga, None, None, None, None, None


class _SegLossWfMeanFn(torch.autograd.Function):
    """Mean of the w^F-weighted (B,B,H,W) loss tensor and its gradient w.r.t. both prediction maps, on the fused kernels."""

    @staticmethod
    def forward(ctx, p_main, p_aux, target, sdf, alpha, wf_amp, main_w, aux_w):
        ctx.save_for_backward(p_main, p_aux, target, sdf)
        ctx.cfg = (alpha, wf_amp, main_w, aux_w)
        return seg_loss_wf_mean(p_main, p_aux, target, alpha, wf_amp, main_w, aux_w, sdf=sdf).to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, up):
        p_main, p_aux, target, sdf = ctx.saved_tensors
        alpha, wf_amp, main_w, aux_w = ctx.cfg
        pm, pa, g = _f(p_main), _f(p_aux), _f(target)
        b, hw = pm.shape[0], pm[0].numel()
        gm, ga = torch.empty_like(pm), torch.empty_like(pa)
        upf = up.to(torch.float32).contiguous()
        L = _lib.lib()
        n = L.csbsr_seg_loss_wf_workspace_bytes(b, hw)
        ws = _ws(n)
        rc = L.csbsr_seg_loss_wf_grad(pm.data_ptr(), pa.data_ptr(), g.data_ptr(), sdf.data_ptr(), b, hw, C.c_float(alpha),
                                      C.c_float(main_w), C.c_float(aux_w), C.c_float(wf_amp), upf.data_ptr(), gm.data_ptr(),
                                      ga.data_ptr(), ws.data_ptr(), n, _lib.stream_ptr())
        _lib.check(rc, "csbsr_seg_loss_wf_grad")
        _lib.count_launch("csbsr_seg_loss_wf_grad")
        return gm, ga, None, None, None, None, None, None


def seg_loss_train(p_main, p_aux, target, alpha, wf_amp=0.0, main_w=1.0, aux_w=0.4, fused=True):
    """MetaSSLossCalc.calc_ss_loss + JointModelWithLoss.multiple_weight (build_model.py:258-278, 422-438) as an autograd
    node.  wf_amp == 0: per-sample (B,) loss on the fused fwd/bwd kernel.  wf_amp != 0: the reference's out_map=True
    path, whose BCE map (B,1,H,W) + Dice term (B,H,W) broadcast to (B,B,H,W) before the w^F weight
    exp(amp*|p.detach() - g|) multiplies it (loss_functions.py:196-210, 284-345; oriented_weight.py:80-83).  fused=True:
    csbsr_seg_loss_wf_mean / _grad evaluate the mean of that tensor and its gradient in closed form; fused=False forms the
    tensor itself with elementwise torch ops (kept for callers that need the per-element map)."""
    g = _f(target)
    sdf = compute_sdf(g)
    if wf_amp == 0:
        return _SegLossFn.apply(p_main, p_aux, g, sdf, alpha, main_w, aux_w)
    if fused and p_aux is not None:
        # the trainer only ever takes .mean() of the (B,B,H,W) tensor (trainer.py:407): compute that scalar and its gradient
        # in closed form on the device and hand back a broadcast view with the reference's shape and the same mean
        b, _, h, w = p_main.shape
        return _SegLossWfMeanFn.apply(p_main, p_aux, g, sdf, alpha, wf_amp, main_w, aux_w).expand(b, b, h, w)

    def combo(p):
        p = p.clamp(min=1e-8)
        wbce = -(g * torch.log(p + 1e-8) + (1 - g) * torch.log(1 - p + 1e-8)) / 2
        dice = 1.0 / g.numel() - (2 * torch.sum(p * g, dim=1) + 1e-6) / (torch.sum(p.pow(2) + g.pow(2)) + 1e-6)
        return alpha * ((wbce + dice) / 2) + (1 - alpha) * (p * sdf)

    loss = main_w * combo(p_main) + aux_w * combo(p_aux)
    return torch.exp(wf_amp * torch.abs(p_main.detach() - g)) * loss


def kbpn_loss_train(sr, hr, lr, kvec, k_gt, weights=(0.4, 0.4, 0, 2), ksize=21, factor=4):
    """KBPNLoss.forward (sr_loss_functions.py:39-56, Get_pseudo_lr :73-102) as an autograd graph: gradients reach the SR
    image (through the L1 terms, the per-sample blur and the antialiased bicubic resize) and the kernel vector."""
    from ..autograd import blur_per_sample, resize_aa
    k = kvec / kvec.sum(dim=1, keepdim=True)
    plr = resize_aa(blur_per_sample(sr, k, ksize, 1), factor)          # csbsr blur / antialiased-bicubic kernels, fwd + bwd
    kn = k.view(-1, 1, ksize, ksize)
    loss = _KBPNLossFn.apply(sr, hr, plr, lr, kn, k_gt, float(weights[0]), float(weights[1]), float(weights[2]))
    return loss, kn


class _KBPNLossFn(torch.autograd.Function):
    """loss[b] = w_hr mean|sr - hr| + w_lr mean|plr - lr| + w_k mean((k - k_gt)^2): csbsr_sr_loss forward, csbsr_sr_loss_bwd backward
    (the autograd of nn.L1Loss / nn.MSELoss behind loss.backward(), sr_loss_functions.py:39-56)."""

    @staticmethod
    def forward(ctx, sr, hr, plr, lr, kn, k_gt, w_hr, w_lr, w_k):
        f = lambda t: t.detach().contiguous().float()
        s, h, p_, l, k, kg = f(sr), f(hr), f(plr), f(lr), f(kn), f(k_gt)
        b = s.shape[0]
        loss = torch.empty(b, dtype=torch.float32, device=s.device)
        L = _lib.lib()
        n = L.csbsr_sr_loss_workspace_bytes(b)
        ws = torch.empty(max(int(n), 8), dtype=torch.uint8, device=s.device)
        rc = L.csbsr_sr_loss(s.data_ptr(), h.data_ptr(), p_.data_ptr(), l.data_ptr(), k.data_ptr(), kg.data_ptr(), b, s[0].numel(),
                             l[0].numel(), k[0].numel(), C.c_float(w_hr), C.c_float(w_lr), C.c_float(w_k), loss.data_ptr(),
                             ws.data_ptr(), n, _lib.stream_ptr())
        _lib.check(rc, "csbsr_sr_loss")
        _lib.count_launch("csbsr_sr_loss")
        ctx.save_for_backward(s, h, p_, l, k, kg)
        ctx.w = (w_hr, w_lr, w_k)
        return loss

    @staticmethod
    def backward(ctx, g):
        s, h, p_, l, k, kg = ctx.saved_tensors
        w_hr, w_lr, w_k = ctx.w
        g = g.contiguous().float()
        d_sr, d_plr = torch.empty_like(s), torch.empty_like(p_)
        d_k = torch.empty_like(k) if (w_k != 0 and ctx.needs_input_grad[4]) else None
        rc = _lib.lib().csbsr_sr_loss_bwd(s.data_ptr(), h.data_ptr(), p_.data_ptr(), l.data_ptr(), k.data_ptr(), kg.data_ptr(),
                                          g.data_ptr(), s.shape[0], s[0].numel(), l[0].numel(), k[0].numel(), C.c_float(w_hr),
                                          C.c_float(w_lr), C.c_float(w_k), d_sr.data_ptr(), d_plr.data_ptr(),
                                          d_k.data_ptr() if d_k is not None else None, _lib.stream_ptr())
        _lib.check(rc, "csbsr_sr_loss_bwd")
        _lib.count_launch("csbsr_sr_loss_bwd")
        return d_sr, None, d_plr, None, d_k, None, None, None, None
