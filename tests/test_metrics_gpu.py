"""Bit-exact parity of the CUDA AIU / Hausdorff sweep (C-ABI csbsr_seg_metrics) with the oracle and with the
golden vectors produced by the unmodified reference; plus the degradation kernels (tolerance 2e-6)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _case(B, H, W, seed, noise=0.08):
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    probs, masks = [], []
    for b in range(B):
        a, c, w = rng.uniform(-0.6, 0.6), rng.uniform(0.25, 0.75) * H, rng.uniform(1.5, 3.5)
        m = (np.abs(yy - (a * xx + c + 4 * np.sin(xx / rng.uniform(4, 9)))) < w).astype(np.float32)
        base = ndimage.gaussian_filter(np.roll(m, int(rng.integers(-3, 4)), axis=1), 1.2)
        probs.append(np.clip(base * rng.uniform(0.8, 1.3) + noise * rng.standard_normal((H, W)), 0, 1).astype(np.float32))
        masks.append(m)
    return np.stack(probs)[:, None], np.stack(masks)[:, None]


def _run(prob, mask, pct):
    from csbsr_b200.engine.inference import seg_metrics
    return seg_metrics(torch.from_numpy(prob), torch.from_numpy(mask), with_hd=True, percent=pct)


def test_golden_vectors_bit_exact():
    g = np.load(os.path.join(GOLD, "metrics_kat.npz"))
    for case in ("a", "b"):
        for pct in (50, 95):
            r = _run(g[case + "_prob"], g[case + "_mask"], pct)
            assert np.array_equal(r["iou"], g[case + "_iou"])
            assert np.array_equal(r["hd"], g[case + "_hd%d" % pct]), (case, pct)
            assert np.array_equal(r["msd"], g[case + "_msd"]), case


@pytest.mark.parametrize("B,H,W,seed,noise", [(2, 64, 96, 1, 0.08), (1, 50, 37, 2, 0.3), (2, 128, 128, 3, 0.02)])
def test_matches_oracle_bit_exact(B, H, W, seed, noise):
    from oracle import metrics_ref as M
    prob, mask = _case(B, H, W, seed, noise)
    inter, union = M.iou_counts(prob, mask)
    for pct in (50, 95):
        r = _run(prob, mask, pct)
        hd, msd = M.distance_metrics(prob, mask, pct)
        assert np.array_equal(r["inter"], inter) and np.array_equal(r["union"], union)
        assert np.array_equal(r["hd"], hd), np.argwhere(r["hd"] != hd)[:5]
        assert np.array_equal(r["msd"], msd), np.argwhere(r["msd"] != msd)[:5]


def test_large_lists_and_sequential_replay_agree():
    """Salt-and-pepper probabilities: tens of thousands of border corners per list (32768-key shared-memory sort and
    the global-memory fallback), checked against the oracle; the block-parallel replay must equal the literal
    sequential one (CSBSR_METRICS_SEQUENTIAL=1)."""
    from oracle import metrics_ref as M
    rng = np.random.default_rng(9)
    H, W = 200, 216
    prob = rng.random((1, 1, H, W)).astype(np.float32)
    mask = (rng.random((1, 1, H, W)) < 0.3).astype(np.float32)
    r = _run(prob, mask, 50)
    os.environ["CSBSR_METRICS_SEQUENTIAL"] = "1"
    try:
        r_seq = _run(prob, mask, 50)
    finally:
        del os.environ["CSBSR_METRICS_SEQUENTIAL"]
    assert np.array_equal(r["hd"], r_seq["hd"]) and np.array_equal(r["msd"], r_seq["msd"])
    sel = [0, 24, 49, 74, 98]
    hd, msd = M.distance_metrics(prob, mask, 50)
    assert np.array_equal(r["hd"][:, sel], hd[:, sel]) and np.array_equal(r["msd"][:, sel], msd[:, sel])
    assert np.array_equal(r["hd"], hd) and np.array_equal(r["msd"], msd)
    prob2, mask2 = _case(2, 96, 128, 5, 0.1)
    a = _run(prob2, mask2, 95)
    os.environ["CSBSR_METRICS_SEQUENTIAL"] = "1"
    try:
        b = _run(prob2, mask2, 95)
    finally:
        del os.environ["CSBSR_METRICS_SEQUENTIAL"]
    assert np.array_equal(a["hd"], b["hd"]) and np.array_equal(a["msd"], b["msd"])


def test_edge_cases():
    from oracle import metrics_ref as M
    H, W = 24, 40
    prob = np.zeros((4, 1, H, W), np.float32)
    mask = np.zeros((4, 1, H, W), np.float32)
    mask[1, 0, 5:9, 3:20] = 1                          # gt only
    prob[2, 0, 5:9, 3:20] = 0.995                      # prediction only, above every threshold
    prob[3] = 1.0; mask[3] = 1.0                       # full image vs full image: borders at the image edge
    thr = M.THRESHOLDS
    prob[0, 0, 0, :10] = thr[:10]                      # exactly on a threshold: p - t > 0 is false
    prob[0, 0, 1, :10] = np.nextafter(thr[:10], np.float32(1))
    mask[0, 0, 0:2, :10] = 0.4                         # below the IoU binarisation, non-zero for the HD one
    r = _run(prob, mask, 50)
    inter, union = M.iou_counts(prob, mask)
    hd, msd = M.distance_metrics(prob, mask, 50)
    assert np.array_equal(r["inter"], inter) and np.array_equal(r["union"], union)
    assert np.array_equal(r["hd"], hd) and np.array_equal(r["msd"], msd)
    assert (r["hd"][1] == W).all() and (r["hd"][2] == W).all() and (r["hd"][3] == 0).all()


def test_full_size_properties():
    """448x448 (BASELINE size): identity and symmetry properties that do not need the (slow) CPU oracle."""
    prob, mask = _case(2, 448, 448, 7, 0.05)
    r = _run(prob, mask, 50)
    assert (r["union"] >= r["inter"]).all() and (np.diff(r["inter"], axis=1) <= 0).all()
    pm = (prob > 0.5).astype(np.float32)
    r2 = _run(pm * 0.75, pm, 50)                        # prediction == gt for thresholds below 0.75
    assert (r2["hd"][:, :74] == 0).all() and (r2["msd"][:, :74] == 0).all() and (r2["iou"][:, :74] == 1.0).all()


@pytest.mark.parametrize("pct", [50, 95])
def test_full_size_448_bit_exact_vs_oracle(pct):
    """One 448x448 noisy-crack image (the BASELINE metric size): AIU counts, HD and MSD of all 99 thresholds are
    bit-identical to the oracle (reference inference.py:293-336 via oracle/metrics_ref.py)."""
    from oracle import metrics_ref as M
    prob, mask = _case(1, 448, 448, 7, 0.05)
    r = _run(prob, mask, pct)
    inter, union = M.iou_counts(prob, mask)
    hd, msd = M.distance_metrics(prob, mask, pct)
    assert np.array_equal(r["inter"], inter) and np.array_equal(r["union"], union)
    assert np.array_equal(r["hd"], hd), np.argwhere(r["hd"] != hd)[:5]
    assert np.array_equal(r["msd"], msd), np.argwhere(r["msd"] != msd)[:5]


def test_degrade_matches_golden_and_oracle():
    from csbsr_b200.data import degrade as G
    g = np.load(os.path.join(GOLD, "degrade.npz"))
    prm = np.concatenate([g["theta"][:, None], g["sigma"]], 1)
    lr, ks, bl = G.degrade(torch.from_numpy(g["hr"]), prm, return_blurred=True)
    assert np.abs(ks.cpu().numpy() - g["kernels"]).max() <= 2e-8
    assert np.abs(bl.cpu().numpy() - g["blur"]).max() <= 2e-6
    assert np.abs(lr.cpu().numpy() - g["lr"]).max() <= 2e-6
    # reference-named entry points
    torch.manual_seed(3); np.random.seed(4)
    k = G.set_blur(21, mode="gaus", isotropic=False)
    torch.manual_seed(3); np.random.seed(4)
    from oracle import degrade_ref as D
    p = G.draw_gaussian_params(1)
    assert np.abs(k.cpu().numpy() - D.make_kernel(*p[0]).numpy()).max() <= 2e-8
    img = torch.from_numpy(g["hr"][0])
    b = G.conv_kernel2d(img, k)
    assert np.abs(b.cpu().numpy() - D.blur(img, k.cpu()).numpy()).max() <= 2e-6
    assert np.abs(G.FactorResize(4, "bicubic")(b).cpu().numpy() - D.downsample(b.cpu()).numpy()).max() <= 2e-6


@pytest.mark.parametrize("shape", [(6, 3, 64, 96), (2, 3, 448, 448), (3, 1, 16, 20), (1, 3, 100, 228)])
def test_fused_degrade_matches_golden_and_oracle(shape):
    """csbsr_degrade_fused (composed 36x36/s4 kernels, 25 border classes) against the reference fixture and the two-step oracle:
    2e-6 abs on the LR image (values in [0,1]), kernels 2e-8 -- the same bounds as the three-launch form."""
    from csbsr_b200.data import degrade as G
    from oracle import degrade_ref as D
    if shape == (6, 3, 64, 96):
        g = np.load(os.path.join(GOLD, "degrade.npz"))
        hr, prm = torch.from_numpy(g["hr"]), np.concatenate([g["theta"][:, None], g["sigma"]], 1)
        want_lr, want_k = g["lr"], g["kernels"]
    else:
        rng = np.random.default_rng(shape[2])
        hr = torch.from_numpy(rng.random(shape).astype(np.float32))
        prm = np.stack([rng.uniform(0, np.pi, shape[0]), rng.uniform(0.2, 4, shape[0]), rng.uniform(0.2, 4, shape[0])], 1)
        lr_o, k_o, _ = D.degrade(hr.expand(-1, 3, -1, -1) if shape[1] == 1 else hr, prm)
        want_lr, want_k = lr_o.numpy()[:, :shape[1]], k_o.numpy()
    lr, ks = G.degrade(hr, prm)
    lr3, ks3, _ = G.degrade(hr, prm, return_blurred=True)              # three-launch form
    err = np.abs(lr.cpu().numpy() - want_lr)
    print("fused degrade", shape, "max err", err.max(), "border ring max", max(err[..., :2, :].max(), err[..., -2:, :].max(),
          err[..., :, :2].max(), err[..., :, -2:].max()), "vs 3-launch", (lr - lr3).abs().max().item())
    assert np.abs(ks.cpu().numpy() - want_k).max() <= 2e-8
    assert err.max() <= 2e-6
    assert (lr - lr3).abs().max().item() <= 2e-6


def test_philox_params_bit_equal_to_oracle():
    from csbsr_b200.data import degrade as G
    from oracle import degrade_ref as D
    for n, seed, off in [(1, 0, 0), (64, 1121, 0), (1000, 2**40 + 3, 2**33 + 5)]:
        got = G.philox_params(n, seed=seed, offset=off).cpu().numpy()
        assert np.array_equal(got, D.philox_params(n, seed=seed, offset=off))


def test_psnr_ssim_kernel_vs_reference_golden_and_oracle():
    """csbsr_psnr_ssim against the reference's PSNR / SSIM outputs and, on a 448^2 batch, the oracle (fp32: 2e-4 dB / 2e-5)."""
    import os
    from csbsr_b200.engine import inference as E
    from oracle import metrics_ref as M
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "psnr_ssim.npz"))
    ps, ss = E.psnr_ssim(torch.from_numpy(g["a"]), torch.from_numpy(g["b"]))
    assert np.abs(ps - g["psnr"]).max() <= 2e-4 and np.abs(ss - g["ssim"]).max() <= 2e-5
    gen = torch.Generator().manual_seed(3)
    a = torch.rand(2, 3, 448, 448, generator=gen)
    b = (a + 0.02 * torch.randn(2, 3, 448, 448, generator=gen)).clamp(0, 1)
    ps, ss = E.psnr_ssim(a, b)
    assert np.abs(ps - M.psnr(a, b)).max() <= 2e-4 and np.abs(ss - M.ssim(a, b)).max() <= 2e-5
    assert np.allclose(E.SSIM()(a, a), 1.0, atol=1e-6)
