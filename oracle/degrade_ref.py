"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's on-the-fly degradation.

Pinned against the UNMODIFIED reference by tests/golden/degrade.npz (tests/golden/gen_golden.py).
  make_kernel    GaussianBlur.make           model/data/blur/blur.py:128-168 (numpy fp64 -> fp32)
  blur           conv_kernel2d               model/data/blur/blur.py:182-200 (depthwise F.conv2d, zero pad)
  downsample     FactorResize('bicubic')     model/data/transforms/transforms.py:516-531 (antialiased bicubic)
"""
import numpy as np
import torch
import torch.nn.functional as F


def make_kernel(theta, sx, sy, size=21):
    r = int(int(size / 2))
    rng = np.linspace(-r, r, size).reshape((1, -1))
    xs = np.tile(rng, (size, 1))
    ys = np.tile(rng.T, (1, size))
    ct, st = np.cos(theta), np.sin(theta)
    sx2, sy2 = 2.0 * (sx ** 2), 2.0 * (sy ** 2)
    a = ct ** 2 / sx2 + st ** 2 / sy2
    b = st * ct * (1.0 / sy2 - 1.0 / sx2)
    c = st ** 2 / sx2 + ct ** 2 / sy2
    k = np.exp(-(a * (xs ** 2) + 2.0 * b * xs * ys + c * (ys ** 2)))
    k = k / k.sum()
    return torch.FloatTensor(k)


def blur(img, kernel):
    """img [C,H,W] fp32, kernel [k,k] fp32."""
    c = img.shape[0]
    k = kernel.shape[-1]
    w = kernel.view(1, 1, k, k).repeat(c, 1, 1, 1)
    return F.conv2d(img.unsqueeze(0), w, stride=1, padding=(k - 1) // 2, groups=c)[0]


def downsample(img, factor=4):
    h, w = img.shape[-2:]
    x = img if img.dim() == 4 else img.unsqueeze(0)
    y = F.interpolate(x, size=(int(h / factor), int(w / factor)), mode="bicubic", antialias=True, align_corners=False)
    return y if img.dim() == 4 else y[0]


def degrade(hr, params, size=21, factor=4):
    """hr [B,3,H,W] fp32, params [B,3] float64 -> (lr, kernels, blurred)."""
    ks, bl, lr = [], [], []
    for i in range(hr.shape[0]):
        k = make_kernel(*[float(v) for v in params[i]], size=size)
        b = blur(hr[i], k)
        ks.append(k); bl.append(b); lr.append(downsample(b, factor))
    return torch.stack(lr), torch.stack(ks), torch.stack(bl)
