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


def gen_alpha_schedule():
    """alpha of BoundaryComboLoss over 40 update_alpha() calls (loss_functions.py:26-41, 76-81) for two resume points."""
    rh.setup()
    import contextlib, io
    from model.utils.loss_functions import BoundaryComboLoss
    out = {}
    for tag, (per_epoch, resume, ratio) in {"a": (7, 0, 1.0), "b": (5, 23, 2.0)}.items():
        with contextlib.redirect_stdout(io.StringIO()):
            fn = BoundaryComboLoss(per_epoch=per_epoch, resume_iter=resume, decrease_ratio=ratio)
        seq = [fn.alpha]
        for i in range(40):
            if i == 20:
                fn.fix_alpha = True
            if i == 26:
                fn.fix_alpha = False
            fn.update_alpha()
            seq.append(fn.alpha)
        out["alpha_" + tag] = np.array(seq, dtype=np.float64)
        out["cfg_" + tag] = np.array([per_epoch, resume, ratio], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "alpha_schedule.npz"), **out)
    print("alpha_schedule.npz", out["alpha_a"][:10], out["alpha_b"][:10])


def gen_psnr_ssim():
    """PSNR / SSIM classes of the reference (model/utils/estimate_metrics.py) on seeded images."""
    rh.setup()
    from model.utils.estimate_metrics import PSNR, SSIM
    g = torch.Generator().manual_seed(51)
    a = torch.rand(3, 3, 40, 56, generator=g)
    b = (a + 0.05 * torch.randn(3, 3, 40, 56, generator=g)).clamp(0, 1)
    b[2] = torch.nn.functional.avg_pool2d(a[2:3], 5, 1, 2)[0]
    np.savez_compressed(os.path.join(HERE, "psnr_ssim.npz"), a=a.numpy(), b=b.numpy(), psnr=PSNR()(a, b), ssim=SSIM()(a, b))
    print("psnr_ssim.npz", PSNR()(a, b), SSIM()(a, b))


B8_DAMP = 0.1


def gen_train(bn_eval=False, hrnet=False, iteration=40000, blurskip=False, batch8=False):
    """One JointModelWithLoss forward + backward of the UNMODIFIED reference at iteration 40000 (all phases active,
    w^F on, m^F = 1), Dropout2d disabled (p = 0) so the step is deterministic: losses and a sample of gradients.
    bn_eval=True additionally puts the BatchNorm layers in eval mode (running statistics): with random weights and a
    batch of 2, batch-statistics BN makes the gradients chaotic under bf16 rounding, so the GPU gradient-parity test
    uses this well-conditioned variant and the batch-statistics variant checks losses / gradient norms."""
    rh.setup()
    import contextlib, io
    from csbsr_b200.modeling import params as P
    from model.modeling.build_model import JointModelWithLoss
    from model.data.transforms.transforms import FactorResize
    from model.engine.trainer import calc_loss
    cfg = rh.make_cfg(wf_amp=1.0, detector="HRNet_OCR" if hrnet else "PSPNet_BlurSkip" if blurskip else "PSPNet")
    if hrnet:
        cfg.SOLVER.TASK_LOSS_WEIGHT = 0.9                       # config #4 (beta = 0.9)
        rh.patch_hrnet_configer()
    with contextlib.redirect_stdout(io.StringIO()):
        m = JointModelWithLoss(cfg, num_train_ds=100, resume_iter=iteration, sr_transforms=FactorResize(4, "bicubic"))
    sd = P.synth_state_dict(P.kbpn_param_shapes(), prefix="sr_model.")
    seg_shapes = P.hrnet_ocr_param_shapes() if hrnet else P.pspnet_param_shapes(blur_dim=441 if blurskip else None)
    sd.update(P.synth_state_dict(seg_shapes, prefix="segmentation_model."))
    if batch8:
        # random-init ResNet with batch-statistics BatchNorm is chaotic under bf16 operand rounding at ANY batch size
        # (tests/tools/conditioning_probe.py: the fp32 oracle with bf16-rounded conv operands moves the segmentation map by
        # 0.08 mean-abs and gradient norms by up to 3.2x on the un-damped weights).  Damping the residual branches
        # (bn2.weight x 0.1, what zero-init-residual training starts from) makes the step well conditioned (probe: every
        # sampled gradient cosine >= 0.975, norms within 8 %) while BatchNorm still runs on batch statistics everywhere.
        for k in sd:
            if ".feats.layer" in k and k.endswith("bn2.weight"):
                sd[k] = sd[k] * B8_DAMP
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("sr_loss_fn") or "vgg" in k.lower() for k in missing), (missing, unexpected)
    m.train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout2d):
            mod.p = 0.0
        if bn_eval and isinstance(mod, torch.nn.modules.batchnorm._BatchNorm):
            mod.eval()
    alpha = 0.63
    m.ss_loss_fn.alpha = alpha
    m.ss_loss_fn.fix_alpha = True
    # batch8: batch-statistics BatchNorm on 8 x 128^2 crops -- statistics over 8 x 16^2 .. 8 x 64^2 values per channel
    # are well conditioned, unlike the batch of 2
    nb, hh, ww = (8, 128, 128) if batch8 else (2, 64, 96)
    _, mask_np = synthetic_case(nb, hh, ww, 77)
    rng = np.random.default_rng(79)
    hr_np = np.clip(0.55 + 0.1 * rng.standard_normal((nb, 3, hh, ww)) - 0.3 * mask_np, 0, 1).astype(np.float32)
    hr = torch.from_numpy(hr_np)
    mask = torch.from_numpy(mask_np)
    g = torch.Generator().manual_seed(78)
    lr = torch.nn.functional.interpolate(hr, size=(hh // 4, ww // 4), mode="bicubic", antialias=True).clamp(0, 1)
    kgt = torch.rand(nb, 1, 21, 21, generator=g)
    kgt = kgt / kgt.sum(dim=(2, 3), keepdim=True)
    with contextlib.redirect_stdout(io.StringIO()):
        seg_loss, sr_loss, seg, sr, kp = m(iteration, lr.clone(), sr_targets=hr.clone(), segment_targets=mask.clone(),
                                           kernel_targets=kgt.clone())

        class A:
            pass
        loss, _, _ = calc_loss(seg_loss, 0.0, sr_loss, 0.0, iteration, cfg, A())
    loss.backward()
    out = {"hr": hr_np, "mask": mask_np, "lr": lr.numpy(), "kgt": kgt.numpy(), "alpha": np.float32(alpha),
           "loss": np.float64(loss.item()), "seg_loss_mean": np.float64(seg_loss.mean().item()),
           "sr_loss": sr_loss.detach().numpy(), "seg_loss_shape": np.array(seg_loss.shape),
           "sr": sr.detach().numpy().astype(np.float16), "seg": seg.detach().numpy().astype(np.float16)}
    if batch8:
        out["bn2_damp"] = np.float32(B8_DAMP)
    names = ["sr_model.feat.0.weight", "sr_model.predictor.feat_ext.2.layer.weight",
             "sr_model.back_projection_stages.0.up.up_conv1.layer.weight",
             "sr_model.back_projection_stages.1.sft.SFT_scale_conv0.weight",
             "sr_model.back_projection_stages.2.kb.kernel_predictor.fe_cat.2.layer.weight",
             "sr_model.back_projection_stages.3.kb.sr_reconst.layer.weight", "sr_model.output_conv.layer.weight",
             "segmentation_model.feats.conv1.weight", "segmentation_model.feats.layer3.2.conv1.weight",
             "segmentation_model.feats.layer4.2.bn2.weight", "segmentation_model.psp.bottleneck.weight",
             "segmentation_model.up_2.conv.0.weight", "segmentation_model.final.0.weight", "segmentation_model.aux.4.bias"]
    if blurskip:
        names = ["segmentation_model.blur_skip.0.conv_scale.0.layer.weight", "segmentation_model.blur_skip.0.conv_shift.1.layer.weight",
                 "segmentation_model.blur_skip.0.conv_scale.0.act.weight", "segmentation_model.blur_skip.1.layer.weight",
                 "segmentation_model.blur_skip.1.norm.weight", "segmentation_model.blur_skip.2.conv_shift.0.layer.bias",
                 "segmentation_model.blur_skip.3.layer.weight"]
    if hrnet:
        names = [n for n in names if n.startswith("sr_model.")] + [
            "segmentation_model.backbone.conv1.weight", "segmentation_model.backbone.layer1.2.conv2.weight",
            "segmentation_model.backbone.stage3.1.branches.2.1.conv1.weight",
            "segmentation_model.backbone.stage4.2.fuse_layers.0.3.0.weight",
            "segmentation_model.backbone.stage4.0.fuse_layers.3.0.1.0.weight",
            "segmentation_model.conv3x3.0.weight", "segmentation_model.aux_head.2.weight",
            "segmentation_model.ocr_distri_head.object_context_block.f_down.0.weight",
            "segmentation_model.ocr_distri_head.conv_bn_dropout.0.weight", "segmentation_model.cls_head.weight"]
    params = dict(m.named_parameters())
    norms = {}
    for k, p_ in params.items():
        if p_.grad is not None:
            norms[k] = float(p_.grad.double().norm().item())
    out["grad_norm_names"] = np.array(sorted(norms))
    out["grad_norms"] = np.array([norms[k] for k in sorted(norms)])
    out["iteration"] = np.int64(iteration)
    out["requires_grad_names"] = np.array(sorted(k for k, p_ in params.items() if p_.requires_grad))
    names = [k for k in names if params[k].grad is not None]
    for k in names:                      # flattened gradients, subsampled with a fixed stride to keep the fixture small
        gflat = params[k].grad.numpy().astype(np.float32).reshape(-1)
        stride = max(1, gflat.size // 20000)
        out["grad:" + k] = gflat[::stride].astype(np.float16 if False else np.float32)
        out["stride:" + k] = np.int64(stride)
    fname = "train_step_hrnet.npz" if hrnet else "train_step_blurskip.npz" if blurskip else \
        "train_step_bneval.npz" if bn_eval else "train_step_b8.npz" if batch8 else "train_step.npz"
    if iteration != 40000:
        fname = "train_step_it%d.npz" % iteration
    np.savez_compressed(os.path.join(HERE, fname), **out)
    print("train_step.npz loss", loss.item(), "seg", seg_loss.mean().item(), "sr", sr_loss.detach().numpy(), "grads", len(norms))


if __name__ == "__main__" and "psnr" in sys.argv[1:]:
    gen_psnr_ssim()

if __name__ == "__main__" and "alpha" in sys.argv[1:]:
    gen_alpha_schedule()

if __name__ == "__main__" and "train" in sys.argv[1:]:
    if "blurskip" in sys.argv[1:]:
        gen_train(bn_eval=True, blurskip=True)
    elif "pretrain" in sys.argv[1:]:
        for it in (5, 15000, 20000, 25000):    # SR-module / kernel-module pre-training, its last iteration, SR-only phase
            gen_train(bn_eval=True, iteration=it)
    elif "hrnet" in sys.argv[1:]:
        gen_train(bn_eval=True, hrnet=True)
    elif "b8" in sys.argv[1:]:
        gen_train(batch8=True)
    else:
        gen_train()
        gen_train(bn_eval=True)
