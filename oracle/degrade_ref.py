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


def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11), the published
    algorithm restated in numpy: counter uint32 [n,4], key uint32 [2] -> uint32 [n,4].  Checker of the throughput-mode draws
    (csbsr_degrade_params_philox); known-answer vectors of the Random123 distribution are asserted in tests/test_oracle_cpu.py."""
    c = np.array(counter, dtype=np.uint64).reshape(-1, 4).copy()
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[:, 0], M1 * c[:, 2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = np.stack([hi1 ^ c[:, 1] ^ k0, lo1, hi0 ^ c[:, 3] ^ k1, lo0], axis=1)
        k0 = (k0 + np.uint64(0x9E3779B9)) & MASK
        k1 = (k1 + np.uint64(0xBB67AE85)) & MASK
    return c.astype(np.uint32)


def philox_params(n, seed=1121, offset=0, range_theta=(0, 180), range_sigma=(0.2, 4)):
    """(theta [rad], sigma_x, sigma_y) float64 [n,3] exactly as csbsr_degrade_params_philox draws them."""
    idx = np.arange(n, dtype=np.uint64) + np.uint64(offset)
    ctr = np.stack([idx & np.uint64(0xFFFFFFFF), idx >> np.uint64(32), np.zeros_like(idx), np.zeros_like(idx)], axis=1)
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, seed >> 32)).astype(np.float64) * (1.0 / 4294967296.0)
    tl, th = range_theta[0] * np.pi / 180, range_theta[1] * np.pi / 180
    return np.stack([tl + (th - tl) * r[:, 0], range_sigma[0] + (range_sigma[1] - range_sigma[0]) * r[:, 1],
                     range_sigma[0] + (range_sigma[1] - range_sigma[0]) * r[:, 2]], axis=1)
