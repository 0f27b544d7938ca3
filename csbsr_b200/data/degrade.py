"""On-the-fly degradation with the reference's function names (model/data/blur/blur.py, transforms.py) running
on the device through the C-ABI: set_blur / GaussianBlur.make -> csbsr_blur_kernel_synth, conv_kernel2d ->
csbsr_blur_per_sample, FactorResize -> csbsr_resize_bicubic_aa, and the batched `degrade` used by the eval loop.

Random draws stay on the host exactly as in the reference (theta from torch.rand, sigmas from np.random.rand:
blur.py:129,170-179), so seeding torch / numpy reproduces the reference's kernels."""
import ctypes as C

import numpy as np
import torch

from .. import _lib


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


def _params_tensor(params, device):
    if torch.is_tensor(params):
        return params.to(device=device, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.asarray(params, dtype=np.float64)).to(device).contiguous()


def draw_gaussian_params(n=1, range_theta=(0, 180), range_sigma=(0.2, 4), range_sigma2=None, isotropic=False):
    """The reference's per-sample draws: returns float64 [n,3] = (theta [rad], sigma_x, sigma_y)."""
    out = np.empty((n, 3), dtype=np.float64)
    for i in range(n):
        theta = ((range_theta[1] - range_theta[0]) * torch.rand(1).item() + range_theta[0]) * np.pi / 180
        a, b = range_sigma
        sx = (b - a) * np.random.rand() + a
        if range_sigma2 is not None:
            a, b = range_sigma2
        sy = (b - a) * np.random.rand() + a
        out[i] = (theta, sx, sx if isotropic else sy)
    return out


def gaussian_kernels(params, size=21):
    """params float64 [n,3] -> fp32 [n,size,size] on the device (GaussianBlur.make, blur.py:128-168)."""
    p = _params_tensor(params, _dev())
    out = torch.empty((p.shape[0], size, size), dtype=torch.float32, device=p.device)
    _lib.check(_lib.lib().csbsr_blur_kernel_synth(p.data_ptr(), out.data_ptr(), p.shape[0], size, _lib.stream_ptr()),
               "csbsr_blur_kernel_synth")
    return out


def set_blur(size=21, device="cuda", mode="gaus", range_gaus_deterioration_ratio=(0.2, 4),
             range_gaus_deterioration_ratio2=None, isotropic=True, **_unused):
    """blur.py:207-238 for mode='gaus' (the only mode the CSBSR configs use, crack_dataset.py:52)."""
    if mode != "gaus":
        raise NotImplementedError("set_blur(mode=%r): only 'gaus' is on the CSBSR path" % (mode,))
    prm = draw_gaussian_params(1, range_sigma=range_gaus_deterioration_ratio,
                               range_sigma2=range_gaus_deterioration_ratio2, isotropic=isotropic)
    return gaussian_kernels(prm, size)[0]


def conv_kernel2d(img, kernel, device="cuda", add_minibatch=True):
    """blur.py:182-200: same kernel for every channel, zero padding, stride 1. img [C,H,W] (or [B,C,H,W] with
    kernel [B,k,k])."""
    x = img.to(device=_dev(), dtype=torch.float32)
    single = x.dim() == 3
    if single:
        x = x.unsqueeze(0)
    x = x.contiguous()
    k = kernel.to(device=_dev(), dtype=torch.float32).reshape(-1, kernel.shape[-2], kernel.shape[-1]).contiguous()
    if k.shape[0] == 1 and x.shape[0] > 1:
        k = k.expand(x.shape[0], -1, -1).contiguous()
    out = torch.empty_like(x)
    n, c, h, w = x.shape
    _lib.check(_lib.lib().csbsr_blur_per_sample(x.data_ptr(), k.data_ptr(), None, out.data_ptr(), n, c, h, w,
                                                k.shape[-1], 1, _lib.stream_ptr()), "csbsr_blur_per_sample")
    return out[0] if single else out


class FactorResize:
    """transforms.py:505-531 with interpolation='bicubic' (antialiased, as torchvision >= 0.17 resizes tensors)."""

    def __init__(self, factor, interpolation="bicubic"):
        if interpolation != "bicubic":
            raise NotImplementedError(interpolation)
        self.factor = factor

    def __call__(self, image, clamp01=False):
        x = image.to(device=_dev(), dtype=torch.float32)
        single = x.dim() == 3
        if single:
            x = x.unsqueeze(0)
        x = x.contiguous()
        n, c, h, w = x.shape
        oh, ow = int(h / self.factor), int(w / self.factor)
        out = torch.empty((n, c, oh, ow), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().csbsr_resize_bicubic_aa(x.data_ptr(), out.data_ptr(), n * c, h, w, oh, ow, int(clamp01),
                                                      _lib.stream_ptr()), "csbsr_resize_bicubic_aa")
        return out[0] if single else out


_ws = {}


def philox_params(n, seed=1121, offset=0, range_theta=(0, 180), range_sigma=(0.2, 4)):
    """Throughput mode of the per-sample draws (blur.py:129,170-179) on the device: float64 [n,3] = (theta [rad], sigma_x,
    sigma_y) from Philox4x32-10 (csbsr_degrade_params_philox); sample i of a run uses counter offset + i."""
    out = torch.empty((n, 3), dtype=torch.float64, device=_dev())
    _lib.check(_lib.lib().csbsr_degrade_params_philox(out.data_ptr(), n, seed, offset, range_theta[0] * np.pi / 180,
                                                      range_theta[1] * np.pi / 180, float(range_sigma[0]), float(range_sigma[1]),
                                                      _lib.stream_ptr()), "csbsr_degrade_params_philox")
    _lib.count_launch("csbsr_degrade_params_philox")
    return out


def degrade(hr, params, ksize=21, factor=4, clamp01=False, return_blurred=False):
    """Batched CrackDataSet.__getitem__ degradation (crack_dataset.py:51-62): hr fp32 [B,3,H,W] + params float64
    [B,3] -> (lr [B,3,H/f,W/f], kernels [B,k,k]).  One fused pass (composed 36x36/s4 kernels, csbsr_degrade_fused) for the
    CSBSR configuration (k = 21, x4); the three-launch form with the intermediate `blurred` image otherwise or on request."""
    x = hr.to(device=_dev(), dtype=torch.float32).contiguous()
    p = _params_tensor(params, x.device)
    b, c, h, w = x.shape
    kernels = torch.empty((b, ksize, ksize), dtype=torch.float32, device=x.device)
    lr = torch.empty((b, c, h // factor, w // factor), dtype=torch.float32, device=x.device)
    L = _lib.lib()
    if not return_blurred and ksize == 21 and factor == 4 and h % 4 == 0 and w % 4 == 0 and h >= 16 and w >= 16:
        need = L.csbsr_degrade_workspace_bytes(b)
        ws = _ws.get((x.device, b))
        if ws is None:
            ws = _ws[(x.device, b)] = torch.empty(need, dtype=torch.uint8, device=x.device)
        _lib.check(L.csbsr_degrade_fused(x.data_ptr(), p.data_ptr(), kernels.data_ptr(), lr.data_ptr(), ws.data_ptr(), need,
                                         b, c, h, w, ksize, factor, int(clamp01), _lib.stream_ptr()), "csbsr_degrade_fused")
        _lib.count_launch("csbsr_degrade")
        return lr, kernels
    blurred = torch.empty_like(x)
    _lib.check(L.csbsr_degrade(x.data_ptr(), p.data_ptr(), kernels.data_ptr(), blurred.data_ptr(),
                               lr.data_ptr(), b, c, h, w, ksize, factor, int(clamp01), _lib.stream_ptr()),
               "csbsr_degrade")
    _lib.count_launch("csbsr_degrade")
    if return_blurred:
        return lr, kernels, blurred
    return lr, kernels
