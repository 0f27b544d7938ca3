"""TEST INFRASTRUCTURE ONLY -- numpy/scipy restatement of the reference's AIU and Hausdorff sweeps.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
Pinned against the UNMODIFIED reference (IoU, calc_distance_metrics, compute_surface_distances) by
tests/golden/gen_golden.py -> tests/golden/metrics_*.npz, and against the known-answer vectors of
SURVEY.md App. E in tests/test_oracle_cpu.py.  Paths below are relative to the reference root.
"""
import math

import numpy as np
from scipy import ndimage

# model/engine/inference.py:49-51 -- torch.Tensor([i*0.01 ...]) is float32
THRESHOLDS = np.array([i * 0.01 for i in range(1, 100)], dtype=np.float64).astype(np.float32)


def binarise(prob):
    """(segment_preds - threshold_map > 0) in float32: inference.py:111.  prob (B,1,H,W) f32 -> (B,99,H,W) bool."""
    p = np.asarray(prob, dtype=np.float32)
    return (p - THRESHOLDS.reshape(1, -1, 1, 1)) > np.float32(0)


def iou_counts(prob, mask):
    """IoU.__call__: model/utils/estimate_metrics.py:72-84 -> integer (intersection, union) per (image, threshold)."""
    pred = binarise(prob)
    tgt = np.asarray(mask, dtype=np.float32) > 0.5
    inter = (pred & tgt).sum(axis=(2, 3))
    union = (pred | tgt).sum(axis=(2, 3))
    return inter.astype(np.int64), union.astype(np.int64)


def iou_from_counts(inter, union, smooth=1e-5):
    return (inter + smooth) / (union + smooth)


def contour_length_table():
    """lookup_tables.create_table_neighbour_code_to_contour_length((1, 1)): lookup_tables.py:330-400."""
    diag = 0.5 * math.sqrt(1 ** 2 + 1 ** 2)
    t = np.zeros(16)
    for code in (0b0001, 0b0010, 0b0100, 0b0111, 0b1000, 0b1011, 0b1101, 0b1110):
        t[code] = diag
    for code in (0b0011, 0b1100):
        t[code] = 1            # horizontal
    for code in (0b0101, 0b1010):
        t[code] = 1            # vertical
    for code in (0b0110, 0b1001):
        t[code] = 2 * diag
    return t


_KERNEL = np.array([[8, 4], [2, 1]])
_TABLE = contour_length_table()


def surface_distances(gt, pred):
    """compute_surface_distances (2-D, spacing (1,1)): surface_distance.py:136-288."""
    both = gt | pred
    if not both.any():
        e = np.array([])
        return e, e, e, e
    ys, xs = np.nonzero(both)
    y0, y1, x0, x1 = ys.min(), ys.max(), xs.min(), xs.max()
    def crop(m):                                      # _crop_to_bounding_box :97-119 (one zero row/col after)
        c = np.zeros((y1 - y0 + 2, x1 - x0 + 2), np.uint8)
        c[:-1, :-1] = m[y0:y1 + 1, x0:x1 + 1]
        return c
    cg, cp = crop(gt), crop(pred)
    ng = ndimage.correlate(cg, _KERNEL, mode="constant", cval=0)
    npd = ndimage.correlate(cp, _KERNEL, mode="constant", cval=0)
    bg = (ng != 0) & (ng != 15)
    bp = (npd != 0) & (npd != 15)
    dg = ndimage.distance_transform_edt(~bg, sampling=(1, 1)) if bg.any() else np.inf * np.ones(bg.shape)
    dp = ndimage.distance_transform_edt(~bp, sampling=(1, 1)) if bp.any() else np.inf * np.ones(bp.shape)
    d_g2p, d_p2g = dp[bg], dg[bp]
    a_g, a_p = _TABLE[ng][bg], _TABLE[npd][bp]
    def srt(d, a):                                    # _sort_distances_surfels :122-133
        s = np.array(sorted(zip(d, a)))
        return s[:, 0], s[:, 1]
    if d_g2p.shape != (0,):
        d_g2p, a_g = srt(d_g2p, a_g)
    if d_p2g.shape != (0,):
        d_p2g, a_p = srt(d_p2g, a_p)
    return d_g2p, d_p2g, a_g, a_p


def robust_hausdorff(d_g2p, d_p2g, a_g, a_p, percent):
    """compute_robust_hausdorff: surface_distance.py:322-359."""
    def one(d, a):
        if len(d) == 0:
            return np.inf
        cum = np.cumsum(a) / np.sum(a)
        idx = np.searchsorted(cum, percent / 100.0)
        return d[min(idx, len(d) - 1)]
    return max(one(d_g2p, a_g), one(d_p2g, a_p))


def distance_metrics(prob, mask, percent=50):
    """calc_distance_metrics: model/engine/inference.py:293-336 -> (hd, msd), each (B, 99) float64."""
    pred_all = binarise(prob)
    B, T = pred_all.shape[:2]
    max_img_len = np.max(pred_all.shape[3:])
    hd = np.zeros((B, T))
    msd = np.zeros((B, T))
    for i in range(B):
        gt = np.asarray(mask[i, 0]).astype(bool)
        for j in range(T):
            d_g2p, d_p2g, a_g, a_p = surface_distances(gt, pred_all[i, j])
            if len(d_g2p) == 0 and len(d_p2g) == 0:
                hd[i, j] = 0
            elif len(d_g2p) == 0 or len(d_p2g) == 0:
                hd[i, j] = max_img_len
            else:
                hd[i, j] = robust_hausdorff(d_g2p, d_p2g, a_g, a_p, percent)
            if np.sum(a_g) == 0 and np.sum(a_p) == 0:
                msd[i, j] = 0
            elif np.sum(a_g) == 0 or np.sum(a_p) == 0:
                msd[i, j] = max_img_len
            else:
                msd[i, j] = (np.sum(d_g2p * a_g) / np.sum(a_g) + np.sum(d_p2g * a_p) / np.sum(a_p)) / 2
    return hd, msd


def psnr(img1, img2):
    """PSNR.__call__ (model/utils/estimate_metrics.py:89-100): 10*log10(1/mse), mse over (C,H,W) per image, fp32 torch."""
    import torch
    a, b = torch.as_tensor(img1).float(), torch.as_tensor(img2).float()
    mse = torch.mean((a - b) ** 2, [1, 2, 3])
    return (10 * torch.log10(1 / mse)).numpy().copy()


def ssim(img1, img2, window_size=11, sigma=1.5):
    """SSIM.forward / _ssim with size_average=False (estimate_metrics.py:134-191): Gaussian window = outer product of the
    normalised 1-D window, depthwise conv with zero padding window_size//2, C1 = 0.01^2, C2 = 0.03^2, mean over (C,H,W)."""
    import math
    import torch
    import torch.nn.functional as F
    a, b = torch.as_tensor(img1).float(), torch.as_tensor(img2).float()
    c = a.shape[1]
    g = torch.Tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    g = (g / g.sum()).unsqueeze(1)
    window = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0).expand(c, 1, window_size, window_size).contiguous()
    conv = lambda t: F.conv2d(t, window, padding=window_size // 2, groups=c)
    mu1, mu2 = conv(a), conv(b)
    mu1_sq, mu2_sq, mu12 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1, s2, s12 = conv(a * a) - mu1_sq, conv(b * b) - mu2_sq, conv(a * b) - mu12
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu12 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return m.mean(1).mean(1).mean(1).numpy().copy()
