"""Generates the committed golden fixtures by running the UNMODIFIED reference (/root/reference) in-process.

Run in the build container only:  python tests/golden/gen_golden.py
Outputs (small .npz files next to this script):
  metrics_kat.npz     inputs + IoU / HD(50,95) / MSD of the reference's IoU + calc_distance_metrics
  degrade.npz         GaussianBlur.make kernels, conv_kernel2d blur and FactorResize outputs
  joint_model.npz     JointModel (KBPN + PSPNet) outputs on csbsr_b200.modeling.params.synth_state_dict weights
  joint_blurskip.npz  the same with DETECTOR_TYPE = PSPNet_BlurSkip
  joint_hrnet.npz     the same with DETECTOR_TYPE = HRNet_OCR (HRNet-W48 + OCR head)
  losses.npz          BoundaryComboLoss (+ w^F map mean), compute_sdf1_1 and KBPNLoss values / gradients
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402


def synthetic_case(B, H, W, seed):
    """Crack-like masks (exactly binary) + noisy probability maps."""
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    masks, probs = [], []
    for b in range(B):
        a, c, w = rng.uniform(-0.6, 0.6), rng.uniform(0.25, 0.75) * H, rng.uniform(1.5, 3.5)
        m = (np.abs(yy - (a * xx + c + 4 * np.sin(xx / rng.uniform(4, 9)))) < w).astype(np.float32)
        base = ndimage.gaussian_filter(np.roll(m, rng.integers(-3, 4), axis=1), 1.2)
        p = np.clip(base * rng.uniform(0.8, 1.3) + 0.08 * rng.standard_normal((H, W)), 0, 1).astype(np.float32)
        masks.append(m)
        probs.append(p)
    return np.stack(probs)[:, None], np.stack(masks)[:, None]


def gen_metrics():
    IoU, calc, sd = rh.metric_fns()
    import model.engine.inference as inf
    th = torch.Tensor([i * 0.01 for i in range(1, 100)]).view(99, 1, 1)
    cases = {}
    prob, mask = synthetic_case(3, 48, 64, 11)
    mask[2] = 0                       # gt empty
    prob2, mask2 = synthetic_case(2, 33, 47, 12)
    prob2[1] = 0.0                    # prediction empty at every threshold
    mask2[0, 0, :, :] = 1.0           # full-image mask: borders only at the image edge
    for name, (p, m) in {"a": (prob, mask), "b": (prob2, mask2)}.items():
        pt, mt = torch.from_numpy(p), torch.from_numpy(m)
        bi = (pt - th > torch.Tensor([0])).float()
        iou = IoU()(bi, mt)
        out = {"prob": p, "mask": m, "iou": iou}
        for pct in (50, 95):
            src = inf.calc_distance_metrics.__code__
            # the reference hard-codes percentile = 50 (inference.py:302); for 95 re-run its own loop body
            if pct == 50:
                hd, msd, _, _ = calc(bi, mt, 0, 0)
            else:
                hd = np.zeros(iou.shape); msd = np.zeros(iou.shape)
                for i in range(p.shape[0]):
                    g = m[i, 0].astype(bool)
                    for j in range(99):
                        s = sd.compute_surface_distances(g, bi[i, j].numpy().astype(bool), spacing_mm=(1, 1))
                        a, b_ = len(s["distances_gt_to_pred"]), len(s["distances_pred_to_gt"])
                        hd[i, j] = 0 if (a == 0 and b_ == 0) else (p.shape[3] if (a == 0 or b_ == 0)
                                                                     else sd.compute_robust_hausdorff(s, pct))
                _, msd, _, _ = calc(bi, mt, 0, 0)
            out["hd%d" % pct] = hd
            out["msd"] = msd
        cases[name] = out
    flat = {"%s_%s" % (n, k): v for n, c in cases.items() for k, v in c.items()}
    np.savez_compressed(os.path.join(HERE, "metrics_kat.npz"), **flat)
    print("metrics_kat.npz", {k: v.shape for k, v in flat.items()})


def gen_degrade():
    GaussianBlur, conv_kernel2d, FactorResize = rh.degrade_fns()
    rng = np.random.default_rng(5)
    out = {}
    thetas, sigmas, kernels = [], [], []
    for i in range(6):
        torch.manual_seed(100 + i)
        np.random.seed(200 + i)
        # replay the reference's own draws: theta from torch.rand, sigma from np.random.rand (blur.py:129,170-179)
        st, sn = torch.get_rng_state(), np.random.get_state()
        k = GaussianBlur(size=21, isotropic=False, range_deterioration_ratio=(0.2, 4)).make().to("cpu")
        torch.set_rng_state(st); np.random.set_state(sn)
        theta = (180 * torch.rand(1).item()) * np.pi / 180
        sx = 3.8 * np.random.rand() + 0.2
        sy = 3.8 * np.random.rand() + 0.2
        thetas.append(theta); sigmas.append((sx, sy)); kernels.append(k.numpy())
    out["theta"] = np.array(thetas); out["sigma"] = np.array(sigmas); out["kernels"] = np.stack(kernels)
    hr = rng.random((6, 3, 64, 96)).astype(np.float32)
    blur, lr = [], []
    fr = FactorResize(4, "bicubic")
    for i in range(6):
        b = conv_kernel2d(torch.from_numpy(hr[i]), torch.from_numpy(kernels[i])).to("cpu")
        blur.append(b.numpy()); lr.append(fr(b).numpy())
    out["hr"] = hr; out["blur"] = np.stack(blur); out["lr"] = np.stack(lr)
    np.savez_compressed(os.path.join(HERE, "degrade.npz"), **out)
    print("degrade.npz", {k: v.shape for k, v in out.items()})


def gen_joint(blur_skip=False, hrnet=False):
    from csbsr_b200.modeling import params as P
    cfg = rh.make_cfg(detector="HRNet_OCR" if hrnet else "PSPNet_BlurSkip" if blur_skip else "PSPNet")
    m = rh.joint_model(cfg)
    sd = P.synth_state_dict(P.kbpn_param_shapes(), prefix="sr_model.")
    if hrnet:
        sd.update(P.synth_state_dict(P.hrnet_ocr_param_shapes(), prefix="segmentation_model."))
    else:
        sd.update(P.synth_state_dict(P.pspnet_param_shapes(blur_dim=441 if blur_skip else None), prefix="segmentation_model."))
    m.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(21)
    x = torch.rand(2, 3, 16, 24, generator=g)
    with torch.no_grad():
        sr, seg, kp = m(x.clone(), torch.zeros(2, 1, 7, 7))
    np.savez_compressed(os.path.join(HERE, "joint_hrnet.npz" if hrnet else "joint_blurskip.npz" if blur_skip else "joint_model.npz"), x=x.numpy(), sr=sr.numpy().astype(np.float16),
                        seg=seg.numpy().astype(np.float16), kp=kp.numpy(),
                        sr_checksum=np.float64(sr.double().sum().item()), seg_checksum=np.float64(seg.double().sum().item()))
    print("joint_model.npz", sr.shape, seg.shape, kp.shape)


def gen_losses():
    """Loss values / gradients of the reference's own classes (BoundaryComboLoss, SegmentFailerOrientedExpWeight, KBPNLoss)."""
    rh.setup()
    import contextlib, io
    from model.utils.loss_functions import BoundaryComboLoss
    from model.utils.oriented_weight import SegmentFailerOrientedExpWeight
    from model.utils.boundary_loss import compute_sdf1_1
    from model.utils.sr_loss_functions import KBPNLoss
    from model.data.transforms.transforms import FactorResize
    cfg = rh.make_cfg()
    _, mask = synthetic_case(3, 40, 56, 31)
    g = torch.from_numpy(mask)
    gen = torch.Generator().manual_seed(3)
    p_main = (0.05 + 0.9 * torch.rand(3, 1, 40, 56, generator=gen))
    p_main[0, 0, :3, :5] = 1e-9                                  # below the clamp
    p_aux = (0.05 + 0.9 * torch.rand(3, 1, 40, 56, generator=gen))
    out = {"mask": mask, "p_main": p_main.numpy(), "p_aux": p_aux.numpy()}
    out["sdf"] = compute_sdf1_1(mask, mask.shape).astype(np.float32)
    alpha = 0.37
    for mode, out_map in (("plain", False), ("map", True)):
        with contextlib.redirect_stdout(io.StringIO()):
            fn = BoundaryComboLoss(per_epoch=10, resume_iter=0, pos_weight=[1, 1], loss_weight=[1, 1], decrease_ratio=1.0,
                                   out_map=out_map)
        fn.alpha = alpha
        pm, pa = p_main.clone().requires_grad_(True), p_aux.clone().requires_grad_(True)
        loss = 1.0 * fn(pm, g, iter_cnt=True) + 0.4 * fn(pa, g, iter_cnt=False)          # build_model.py:258-270
        if out_map:
            loss = SegmentFailerOrientedExpWeight(cfg, 1.0, 1.0)(pm, g) * loss           # build_model.py:433-434
            out["map_shape"] = np.array(loss.shape)
            out["map_mean"] = np.float64(loss.double().mean().item())
            out["map_mean_f32"] = np.float32(loss.mean().item())
        else:
            out["plain_loss"] = loss.detach().numpy()
            up = torch.tensor([0.3, 0.5, 0.2])
            (loss * up).sum().backward()
            out["plain_upstream"] = up.numpy()
            out["plain_grad_main"] = pm.grad.numpy()
            out["plain_grad_aux"] = pa.grad.numpy()
    out["alpha"] = np.float32(alpha)
    # KBPNLoss (weights [0.4, 0.4, 0, 2] -> kernel term weight 0)
    sr = torch.rand(2, 3, 48, 64, generator=gen)
    hr = torch.rand(2, 3, 48, 64, generator=gen)
    lr = torch.rand(2, 3, 12, 16, generator=gen)
    kv = torch.rand(2, 441, generator=gen)
    kmap = kv.view(2, 441, 1, 1).expand(2, 441, 12, 16).contiguous()
    kgt = torch.rand(2, 1, 21, 21, generator=gen); kgt = kgt / kgt.sum(dim=(2, 3), keepdim=True)
    with contextlib.redirect_stdout(io.StringIO()):
        kl = KBPNLoss(cfg, FactorResize(4, "bicubic"))
    l, kp = kl(sr, hr, lr, kmap, kgt, None, None, 40000)
    out.update({"sr": sr.numpy(), "hr": hr.numpy(), "lr": lr.numpy(), "kvec": kv.numpy(), "kgt": kgt.numpy(),
                "kbpn_loss": l.numpy(), "kbpn_kernel": kp.numpy()})
    np.savez_compressed(os.path.join(HERE, "losses.npz"), **out)
    print("losses.npz", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    which = sys.argv[1:] or ["metrics", "degrade", "joint"]
    if "metrics" in which:
        gen_metrics()
    if "degrade" in which:
        gen_degrade()
    if "joint" in which:
        gen_joint()
    if "losses" in which or not sys.argv[1:]:
        gen_losses()
    if "blurskip" in which or not sys.argv[1:]:
        gen_joint(blur_skip=True)
    if "hrnet" in which or not sys.argv[1:]:
        gen_joint(hrnet=True)
