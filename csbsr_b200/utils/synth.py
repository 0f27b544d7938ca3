"""Synthetic crack images + exactly-binary masks and degradation parameters: the
benchmark and the parity tests run on these because the crack-segmentation dataset is not available offline."""
import numpy as np
import torch
import torch.nn.functional as F


def crack_image(idx, size=448):
    """-> (hr [3,size,size] fp32 in [0,1], mask [1,size,size] in {0.0, 1.0}); seeded by 1121 + idx (cfg.SEED)."""
    g = torch.Generator().manual_seed(1121 + idx)
    H = W = size
    low = torch.randn(1, 1, H // 8 + 2, W // 8 + 2, generator=g)
    low = F.interpolate(low, size=(H, W), mode="bilinear", align_corners=False)[0, 0]
    tex = 0.55 + 0.08 * low + 0.03 * torch.randn(H, W, generator=g)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    mask = torch.zeros(H, W, dtype=torch.bool)
    img = tex.clone()
    n_lines = int(torch.randint(1, 4, (1,), generator=g))
    for _ in range(n_lines):
        width = float(torch.empty(1).uniform_(2.0, 6.0, generator=g))
        val = float(torch.empty(1).uniform_(0.15, 0.30, generator=g))
        p = torch.empty(2).uniform_(0.1 * size, 0.9 * size, generator=g)
        ang = float(torch.empty(1).uniform_(0, 2 * np.pi, generator=g))
        dist = torch.full((H, W), 1e9)
        for _seg in range(24):
            ang += float(torch.randn(1, generator=g)) * 0.45
            q = p + torch.tensor([np.sin(ang), np.cos(ang)], dtype=torch.float32) * (size / 16.0)
            d = q - p
            t = (((yy - p[0]) * d[0] + (xx - p[1]) * d[1]) / (d * d).sum()).clamp(0, 1)
            dist = torch.minimum(dist, torch.sqrt((yy - (p[0] + t * d[0])) ** 2 + (xx - (p[1] + t * d[1])) ** 2))
            p = q
        line = dist < width / 2
        mask |= line
        img = torch.where(line, torch.full_like(img, val) + 0.02 * torch.randn(H, W, generator=g), img)
    hr = img.clamp(0, 1).unsqueeze(0).repeat(3, 1, 1)
    hr = (hr + 0.01 * torch.randn(3, H, W, generator=g)).clamp(0, 1)
    return hr.contiguous(), mask.float().unsqueeze(0)


def batch(start, n, size=448):
    hrs, masks = zip(*[crack_image(start + i, size) for i in range(n)])
    return torch.stack(hrs), torch.stack(masks)


def degradation_params(n, seed=5):
    """theta ~ U(0, pi), sigma_x, sigma_y ~ U(0.2, 4): float64 [n,3] (seed 5 = make_test_blur.py:85)."""
    rng = np.random.default_rng(seed)
    return np.stack([rng.uniform(0, np.pi, n), rng.uniform(0.2, 4.0, n), rng.uniform(0.2, 4.0, n)], axis=1)


def model_state_dict(seed=1121):
    from ..modeling import params as P
    sd = P.synth_state_dict(P.kbpn_param_shapes(), seed=seed, prefix="sr_model.")
    sd.update(P.synth_state_dict(P.pspnet_param_shapes(), seed=seed, prefix="segmentation_model."))
    return sd
