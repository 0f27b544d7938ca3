"""Training-path kernels: conv weight / data gradients on the tcgen05 engine against
torch autograd in fp32 on the same bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rand_nhwc(n, h, w, c, cpad, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.zeros(n, h, w, cpad, dtype=torch.bfloat16)
    t[..., :c] = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16)
    return t.cuda()


WGRAD_CASES = [
    # (n, h, w, cin, cout, k, stride, pad, dil)
    (2, 20, 24, 64, 64, 3, 1, 1, 1),
    (2, 17, 9, 128, 192, 3, 1, 2, 2),
    (1, 33, 40, 64, 128, 3, 2, 1, 1),
    (2, 16, 16, 256, 64, 1, 1, 0, 1),
    (1, 32, 48, 128, 128, 8, 4, 2, 1),
    (3, 8, 8, 320, 64, 3, 1, 1, 1),
]


@pytest.mark.parametrize("case", WGRAD_CASES)
def test_conv_wgrad_vs_autograd(case):
    from csbsr_b200 import kernels as K
    n, h, w, ci, co, k, st, pad, dil = case
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    x = _rand_nhwc(n, h, w, ci, K.round_up(ci, 64), 1)
    oh = (h + 2 * pad - dil * (k - 1) - 1) // st + 1
    ow = (w + 2 * pad - dil * (k - 1) - 1) // st + 1
    dy = _rand_nhwc(n, oh, ow, co, K.round_up(co, 64), 2)
    taps = [(r * dil - pad, s * dil - pad) for r in range(k) for s in range(k)]
    wg = K.wgrad(K.Fmap(dy), K.Fmap(x), taps, stride=st)
    got = wg[:co, :, :ci].reshape(co, k, k, ci).permute(0, 3, 1, 2)
    wt = torch.zeros(co, ci, k, k, device="cuda", requires_grad=True)
    y = F.conv2d(x[..., :ci].permute(0, 3, 1, 2).float(), wt, stride=st, padding=pad, dilation=dil)
    y.backward(dy[..., :co].permute(0, 3, 1, 2).float())
    ref = wt.grad
    err = (got - ref).abs().max().item()
    print("wgrad", case, "max err", err, "ref max", ref.abs().max().item())
    assert err <= 2e-3 * ref.abs().max().item() + 1e-3


def test_deconv_wgrad_vs_autograd():
    """ConvTranspose2d(8, 4, 2): G = layer input, S = dL/dy."""
    from csbsr_b200 import kernels as K
    torch.backends.cudnn.allow_tf32 = False
    n, h, w, ci, co = 2, 12, 16, 128, 64
    x = _rand_nhwc(n, h, w, ci, 128, 3)
    dy = _rand_nhwc(n, 4 * h, 4 * w, co, 64, 4)
    taps = [(r - 2, s - 2) for r in range(8) for s in range(8)]
    wg = K.wgrad(K.Fmap(x), K.Fmap(dy), taps, stride=4)
    got = wg[:ci, :, :co].reshape(ci, 8, 8, co).permute(0, 3, 1, 2)
    wt = torch.zeros(ci, co, 8, 8, device="cuda", requires_grad=True)
    y = F.conv_transpose2d(x.permute(0, 3, 1, 2).float(), wt, stride=4, padding=2)
    y.backward(dy.permute(0, 3, 1, 2).float())
    err = (got - wt.grad).abs().max().item()
    print("deconv wgrad max err", err, "ref max", wt.grad.abs().max().item())
    assert err <= 2e-3 * wt.grad.abs().max().item() + 1e-3


CONV_CASES = [
    # (n, h, w, cin, cout, k, stride, pad, dil, bias)
    (2, 20, 24, 64, 49, 3, 1, 1, 1, True),
    (2, 24, 16, 128, 128, 3, 1, 2, 2, False),
    (1, 32, 40, 64, 128, 3, 2, 1, 1, False),
    (2, 16, 24, 64, 128, 1, 2, 0, 1, False),
    (1, 32, 48, 128, 128, 8, 4, 2, 1, False),
    (2, 30, 22, 3, 64, 7, 2, 3, 1, False),
    (2, 12, 12, 569, 569, 3, 1, 1, 1, True),
]


def _rel(a, b):
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_autograd_fn(case):
    """y, dx, dw, db of the tcgen05 conv Function against torch autograd in fp32 (bf16-rounded operands)."""
    from csbsr_b200 import autograd as A
    n, h, w, ci, co, k, st, pad, dil, has_bias = case
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(5)
    x0 = torch.randn(n, ci, h, w, generator=g).to(torch.bfloat16).float().cuda()
    wt = (torch.randn(co, ci, k, k, generator=g) * (2.0 / (ci * k * k)) ** 0.5).to(torch.bfloat16).float().cuda()
    b = torch.randn(co, generator=g).cuda() if has_bias else None
    x_ref = x0.clone().requires_grad_(True)
    w_ref = wt.clone().requires_grad_(True)
    b_ref = b.clone().requires_grad_(True) if has_bias else None
    y_ref = F.conv2d(x_ref, w_ref, b_ref, stride=st, padding=pad, dilation=dil)
    up = torch.randn(y_ref.shape, generator=torch.Generator().manual_seed(6)).to(torch.bfloat16).float().cuda()
    y_ref.backward(up)

    x = x0.clone().requires_grad_(True)
    w_ = wt.clone().requires_grad_(True)
    b_ = b.clone().requires_grad_(True) if has_bias else None
    y = A.conv2d(A.to_nhwc(x), w_, b_, stride=st, padding=pad, dilation=dil)
    assert y.shape[3] == A.cpad(co) and (y[..., co:] == 0).all()
    y.backward(A.to_nhwc(up))
    print("conv fn", case, "y", _rel(A.to_nchw(y, co), y_ref), "dx", _rel(x.grad, x_ref.grad), "dw", _rel(w_.grad, w_ref.grad))
    assert _rel(A.to_nchw(y, co).detach(), y_ref.detach()) <= 1e-2
    assert _rel(x.grad, x_ref.grad) <= 1e-2
    assert _rel(w_.grad, w_ref.grad) <= 1e-2
    if has_bias:
        assert _rel(b_.grad, b_ref.grad) <= 1e-2


def test_deconv8s4_autograd_fn():
    from csbsr_b200 import autograd as A
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(7)
    for ci, co in ((128, 128), (3, 128)):
        x0 = torch.randn(2, ci, 10, 12, generator=g).to(torch.bfloat16).float().cuda()
        wt = (torch.randn(ci, co, 8, 8, generator=g) * 0.05).to(torch.bfloat16).float().cuda()
        x_ref, w_ref = x0.clone().requires_grad_(True), wt.clone().requires_grad_(True)
        y_ref = F.conv_transpose2d(x_ref, w_ref, stride=4, padding=2)
        up = torch.randn(y_ref.shape, generator=g).to(torch.bfloat16).float().cuda()
        y_ref.backward(up)
        x, w_ = x0.clone().requires_grad_(True), wt.clone().requires_grad_(True)
        y = A.deconv8s4(A.to_nhwc(x), w_)
        y.backward(A.to_nhwc(up))
        print("deconv fn", ci, co, _rel(A.to_nchw(y, co), y_ref), _rel(x.grad, x_ref.grad), _rel(w_.grad, w_ref.grad))
        assert _rel(A.to_nchw(y, co).detach(), y_ref.detach()) <= 1e-2
        assert _rel(x.grad, x_ref.grad) <= 1e-2 and _rel(w_.grad, w_ref.grad) <= 1e-2


def _train_model(hrnet=False, blurskip=False):
    from csbsr_b200.config import cfg
    from csbsr_b200.modeling.build_model import JointModelWithLoss
    from csbsr_b200.modeling import params as P
    c = cfg.clone()
    c.merge_from_file("config/config_csbsr_pspnet.yaml")
    c.SOLVER.SEG_FAIL_ORIENTED_WEIGHT4SS_AMP = 1.0
    if hrnet:
        c.MODEL.DETECTOR_TYPE = "HRNet_OCR"
        c.SOLVER.TASK_LOSS_WEIGHT = 0.9
    if blurskip:
        c.MODEL.DETECTOR_TYPE = "PSPNet_BlurSkip"
    m = JointModelWithLoss(c, num_train_ds=100, resume_iter=40000)
    sd = P.synth_state_dict(P.kbpn_param_shapes(), prefix="sr_model.")
    sd.update(P.synth_state_dict(P.hrnet_ocr_param_shapes() if hrnet else P.pspnet_param_shapes(blur_dim=441 if blurskip else None),
                                 prefix="segmentation_model."))
    m.load_state_dict(sd, strict=True)
    return m.cuda(), sd, c


@pytest.mark.parametrize("variant", ["pspnet_bneval", "pspnet", "pspnet_b8", "hrnet_bneval", "blurskip_bneval", "it5", "it15000", "it20000", "it25000"])
def test_train_step_vs_reference_golden(variant):
    """One joint training step (iteration 40000, w^F on, Dropout2d off) of the tcgen05 training graph against the
    unmodified reference's losses and gradients (tests/golden/train_step*.npz; the fp32 oracle is pinned to the same
    fixtures on CPU).

    bn_eval=True  (BatchNorm on running statistics): well conditioned -> losses within 2 %, gradient cosine >= 0.98 for
                  every sampled parameter, norms within 10 %.
    pspnet_b8     (batch statistics on 8 x 128^2 crops, residual branches of the ResNet damped: bn2.weight x 0.1 in the fixture
                  and here -- tests/tools/conditioning_probe.py shows the un-damped random-init ResNet chaotic under bf16
                  operand rounding at any batch size, and this one well conditioned: fp32 oracle exact vs bf16-rounded
                  operands cosine >= 0.975, norms within 8 %): losses within 2 %, every sampled gradient cosine >= 0.95 and
                  norm within 15 %, all gradient norms within 15 %.
    bn_eval=False (batch statistics, batch of 2, random weights): the network is chaotic under bf16 rounding -- the fp32
                  oracle with its conv operands rounded to bf16 decorrelates from the exact one just as much (cosine
                  0.2-0.9 below the heads) -- so this variant checks losses (3 %), gradient norms (within a factor of 2, at most 20 % of the tensors off by more than 50 %) and the heads."""
    import os
    import numpy as np
    from csbsr_b200.engine.losses import calc_loss
    # itN: the pre-training phases (5: SR modules with the ground-truth kernel, 15000: kernel predictors only, 20000: its
    # last iteration where KBPN re-enables its SR layers one step early, 25000: whole SR net, loss = sr_loss throughout)
    bn_eval, hrnet = variant.endswith("bneval") or variant.startswith("it"), variant.startswith("hrnet")
    b8 = variant == "pspnet_b8"
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                             {"pspnet": "train_step.npz", "pspnet_b8": "train_step_b8.npz", "pspnet_bneval": "train_step_bneval.npz", "blurskip_bneval": "train_step_blurskip.npz",
                              "hrnet_bneval": "train_step_hrnet.npz"}.get(variant, "train_step_%s.npz" % variant)))
    it = int(g["iteration"]) if "iteration" in g.files else 40000
    m, sd, c = _train_model(hrnet, variant.startswith("blurskip"))    # blurskip: only segmentation_model.blur_skip.* is trained
    if b8:
        with torch.no_grad():
            for k, p_ in m.named_parameters():
                if ".feats.layer" in k and k.endswith("bn2.weight"):
                    p_.mul_(float(g["bn2_damp"]))
    m.train()
    m.dropout = False
    m.freeze_bn = bn_eval
    m.ss_loss_fn.alpha = float(g["alpha"])
    lr, hr, mask, kgt = (torch.from_numpy(g[k]).cuda() for k in ("lr", "hr", "mask", "kgt"))
    for p_ in m.parameters():
        p_.grad = torch.zeros_like(p_)
    seg_loss, sr_loss, seg, sr, kp = m(it, lr, sr_targets=hr, segment_targets=mask, kernel_targets=kgt)
    loss = calc_loss(sr_loss, seg_loss.mean(), c.SOLVER.TASK_LOSS_WEIGHT, it, c)
    loss.backward()
    torch.cuda.synchronize()
    seg_err = np.abs(seg.detach().cpu().numpy() - g["seg"].astype(np.float32)).mean()
    print("train loss", loss.item(), "ref", float(g["loss"]), "seg", seg_loss.mean().item(), float(g["seg_loss_mean"]),
          "sr", sr_loss.tolist(), g["sr_loss"].tolist(), "seg mean abs diff", seg_err)
    tol = 2e-2 if (bn_eval or b8) else 3e-2
    assert tuple(seg_loss.shape) == tuple(g["seg_loss_shape"])
    assert abs(loss.item() - float(g["loss"])) <= tol * abs(float(g["loss"]))
    assert abs(seg_loss.mean().item() - float(g["seg_loss_mean"])) <= tol * abs(float(g["seg_loss_mean"]))
    assert np.abs(sr_loss.detach().cpu().numpy() - g["sr_loss"]).max() <= 2e-2 * g["sr_loss"].max()
    assert np.abs(sr.detach().cpu().numpy() - g["sr"].astype(np.float32)).max() <= 5e-2
    assert seg_err <= (1e-2 if bn_eval else 4e-2 if b8 else 0.15)
    params = dict(m.named_parameters())
    for k in [k[5:] for k in g.files if k.startswith("grad:")]:
        ref = g["grad:" + k].astype(np.float64)
        got = params[k].grad.detach().cpu().numpy().reshape(-1)[::int(g["stride:" + k])].astype(np.float64)
        cos = float((got * ref).sum() / np.sqrt((got * got).sum() * (ref * ref).sum()))
        ratio = np.sqrt((got * got).sum() / (ref * ref).sum())
        print("  grad %-75s cos %.5f  norm ratio %.3f" % (k, cos, ratio))
        if bn_eval:
            assert cos >= 0.98 and abs(ratio - 1) <= 0.1, (k, cos, ratio)
        elif b8:
            assert cos >= 0.95 and abs(ratio - 1) <= 0.15, (k, cos, ratio)
        else:
            assert abs(ratio - 1) <= 1.0, (k, ratio)          # chaotic variant: within a factor of 2
            if k in ("segmentation_model.final.0.weight", "segmentation_model.aux.4.bias"):
                assert cos >= 0.99, (k, cos)
    # every trainable tensor received a finite gradient; whole-model gradient norms against the reference's
    names, norms = list(g["grad_norm_names"]), g["grad_norms"]
    gmax = float(np.max(norms))
    alt = {}
    if b8:
        # second opinion for the ill-conditioned tensors: the SAME pinned fp32 oracle with its conv operands rounded to bf16
        # (tests/tools/conditioning_probe.py --save): scalar PReLU slopes (sum_{x<0} dy*x cancels) and kb.sr_reconst move by
        # 17 % .. 4x under operand rounding alone -- a tensor passes when its norm is within tolerance of the unmodified
        # reference's OR of that run's
        a = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_step_b8_bf16ops.npz"))
        alt = dict(zip([str(n) for n in a["names"]], a["norm_bf16ops"]))
    bad = []
    for k, p_ in params.items():
        assert p_.grad is not None and torch.isfinite(p_.grad).all(), k
        if k not in names:                       # frozen / not on the loss path in this phase: the reference has grad None
            assert p_.grad.abs().max().item() == 0, k
        if k in names:
            ref_n = norms[names.index(k)]
            got_n = p_.grad.double().norm().item()
            if b8:
                if ref_n <= 1e-6 * gmax:        # conv bias in front of a batch-statistics BatchNorm: the exact gradient is 0
                    assert got_n <= 1e-4 * gmax, (k, got_n)
                    continue
                alt_n = float(alt.get(k, ref_n))
                if p_.numel() == 1:
                    # scalar PReLU slope: sum_{x<0} dy*x over ~1e6 signed terms cancelling to ~1e-3 of the gradient scale;
                    # operand rounding alone moves it by up to 4x in the oracle (probe) -> absolute bound
                    # (the oracle's own shift under operand rounding, |alt - ref|, sets the scale of the noise)
                    ok = abs(got_n - ref_n) <= max(0.5 * max(ref_n, alt_n), 2 * abs(alt_n - ref_n), 5e-3)
                else:
                    # kb.sr_reconst (3 output channels feeding two nearly cancelling paths): the oracle moves by 17-18 % under
                    # operand rounding alone (probe) -> 30 %; every other tensor 15 % of either reference
                    tol_k = 0.30 if k.endswith("kb.sr_reconst.layer.weight") else 0.15
                    ok = abs(got_n / ref_n - 1) <= tol_k or abs(got_n / (alt_n + 1e-30) - 1) <= tol_k
                if not ok:
                    bad.append((k, got_n, ref_n, alt_n))
                continue
            if p_.numel() == 1:
                # scalar PReLU slopes: sum_{x<0} dy*x cancels down to 1e-5..1e-3, so bf16 activations move it by O(1)
                # relative (the fp32 oracle with bf16-rounded conv operands shows the same outliers): absolute bound
                if abs(got_n - ref_n) > max(0.5 * ref_n, 2e-3):
                    bad.append((k, p_.grad.item(), ref_n))
                continue
            if ref_n == 0:                      # OCR f_pixel / f_object: softmax over K = 1 object region -> exactly zero
                assert p_.grad.abs().max().item() == 0, k
                continue
            r = got_n / (ref_n + 1e-30)
            if abs(r - 1) > (0.15 if bn_eval else 0.5):
                bad.append((k, r))
    print("grad-norm outliers:", bad[:10], len(bad), "of", len(names))
    assert len(bad) <= (0 if (bn_eval or b8) else len(names) // 5)


def test_prelu_fn_vs_torch():
    from csbsr_b200 import autograd as A
    g = torch.Generator().manual_seed(11)
    x0 = torch.randn(2, 20, 24, 64, generator=g).to(torch.bfloat16).cuda()
    up = torch.randn(2, 20, 24, 64, generator=g).to(torch.bfloat16).cuda()
    a0 = torch.tensor([0.17]).cuda()
    x, a = x0.clone().requires_grad_(True), a0.clone().requires_grad_(True)
    y = A.prelu(x, a)
    y.backward(up)
    xr, ar = x0.float().requires_grad_(True), a0.clone().requires_grad_(True)
    yr = F.prelu(xr, ar)
    yr.backward(up.float())
    assert (y.float() - yr).abs().max().item() <= 2e-2
    assert (x.grad.float() - xr.grad).abs().max().item() <= 2e-2
    assert abs(a.grad.item() - ar.grad.item()) <= 1e-3 * abs(ar.grad.item()) + 1e-3


def test_fused_adam_vs_torch_adam():
    """Five steps of csbsr_adam_step on flat buffers against torch.optim.Adam (fp32, same gradients)."""
    from csbsr_b200.engine.optim import FusedAdam
    g = torch.Generator().manual_seed(12)
    shapes = [(64, 3, 3, 3), (49,), (1,), (128, 64, 3, 3), (5, 7)]
    ps = [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt = FusedAdam(ps, lr=2e-5, lr_lambda=lambda it: 10 if it >= 3 else 1)
    topt = torch.optim.Adam(ref, lr=2e-5, betas=(0.9, 0.999), eps=1e-8)
    sched = torch.optim.lr_scheduler.LambdaLR(topt, lr_lambda=lambda it: 10 if it >= 3 else 1)
    for step in range(7):
        frozen = {1, 3} if step in (2, 3, 4) else set()          # pre-training phases freeze modules: torch skips grad None
        for i, (p, r) in enumerate(zip(ps, ref)):
            p.requires_grad_(i not in frozen)
            if i in frozen:
                r.grad = None
                continue
            gr = torch.randn(p.shape, generator=g).cuda() * 10 ** float(torch.randint(-4, 2, (1,), generator=g))
            p.grad.copy_(gr)
            r.grad = gr.clone()
        opt.step()
        opt.scheduler_step()
        topt.step()
        sched.step()
        for p, r in zip(ps, ref):
            assert (p.grad == 0).all()
            assert torch.allclose(p.detach(), r.detach(), rtol=2e-6, atol=1e-9), (step, (p - r).abs().max().item())


@pytest.mark.parametrize("stride", [1, 4])
def test_blur_fn_vs_torch_autograd(stride):
    """Per-sample blur forward / d input / d kernel against grouped F.conv2d autograd."""
    from csbsr_b200 import autograd as A
    g = torch.Generator().manual_seed(31)
    B, H, W, k = 3, 40, 52, 21
    x0 = torch.rand(B, 3, H, W, generator=g).cuda()
    k0 = torch.rand(B, k * k, generator=g).cuda()
    k0 = k0 / k0.sum(1, keepdim=True)
    up = torch.randn(B, 3, (H - 1) // stride + 1, (W - 1) // stride + 1, generator=g).cuda()
    x, kv = x0.clone().requires_grad_(True), k0.clone().requires_grad_(True)
    y = A.blur_per_sample(x, kv, k, stride)
    y.backward(up)
    xr, kr = x0.clone().requires_grad_(True), k0.clone().requires_grad_(True)
    wgt = kr.view(B, 1, 1, k, k).expand(B, 3, 1, k, k).reshape(B * 3, 1, k, k)
    yr = F.conv2d(xr.reshape(1, B * 3, H, W), wgt, stride=stride, padding=10, groups=B * 3).view(B, 3, *up.shape[2:])
    yr.backward(up)
    assert (y - yr).abs().max().item() <= 2e-6
    assert (x.grad - xr.grad).abs().max().item() <= 1e-5 * max(1.0, xr.grad.abs().max().item())
    assert (kv.grad - kr.grad).abs().max().item() <= 2e-4 * kr.grad.abs().max().item()


def test_resize_aa_fn_vs_torch_autograd():
    from csbsr_b200 import autograd as A
    g = torch.Generator().manual_seed(32)
    x0 = torch.rand(2, 3, 48, 64, generator=g).cuda()
    up = torch.randn(2, 3, 12, 16, generator=g).cuda()
    x = x0.clone().requires_grad_(True)
    y = A.resize_aa(x, 4)
    y.backward(up)
    xr = x0.clone().requires_grad_(True)
    yr = F.interpolate(xr, size=(12, 16), mode="bicubic", antialias=True, align_corners=False)
    yr.backward(up)
    assert (y - yr).abs().max().item() <= 2e-6
    assert (x.grad - xr.grad).abs().max().item() <= 2e-6


@pytest.mark.parametrize("training,relu,with_res,c,pitch", [
    (True, True, True, 64, 64), (True, False, False, 48, 64), (False, True, False, 128, 128), (False, False, True, 256, 256),
    (True, True, False, 1024, 1024),
])
def test_batch_norm_fn_vs_torch(training, relu, with_res, c, pitch):
    """csbsr_bn_stats / _apply / _backward (batch or running statistics, fused residual + ReLU, channel padding) against
    F.batch_norm + add + relu autograd in fp32 on the same bf16-rounded input."""
    from csbsr_b200 import autograd as A
    g = torch.Generator().manual_seed(41)
    n, h, w = 3, 10, 14
    x0 = torch.zeros(n, h, w, pitch, dtype=torch.bfloat16)
    x0[..., :c] = (torch.randn(n, h, w, c, generator=g) * 1.5 + 0.3).to(torch.bfloat16)
    r0 = torch.zeros_like(x0)
    r0[..., :c] = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16)
    up = torch.zeros_like(x0)
    up[..., :c] = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16)
    x0, r0, up = x0.cuda(), r0.cuda(), up.cuda()
    gam0, bet0 = (0.5 + torch.rand(c, generator=g)).cuda(), (0.1 * torch.randn(c, generator=g)).cuda()
    rm0, rv0 = (0.1 * torch.randn(c, generator=g)).cuda(), (0.5 + torch.rand(c, generator=g)).cuda()

    x, res = x0.clone().requires_grad_(True), r0.clone().requires_grad_(True)
    gam, bet = gam0.clone().requires_grad_(True), bet0.clone().requires_grad_(True)
    rm, rv = rm0.clone(), rv0.clone()
    y = A.batch_norm(x, gam, bet, rm, rv, training, 0.1, 1e-5, relu=relu, res=res if with_res else None)
    y.backward(up)

    xr = x0[..., :c].float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    rr = r0[..., :c].float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    gr, br = gam0.clone().requires_grad_(True), bet0.clone().requires_grad_(True)
    rm2, rv2 = rm0.clone(), rv0.clone()
    yr = F.batch_norm(xr, rm2, rv2, gr, br, training=training, momentum=0.1, eps=1e-5)
    if with_res:
        yr = yr + rr
    if relu:
        yr = F.relu(yr)
    yr.backward(up[..., :c].float().permute(0, 3, 1, 2))
    nchw = lambda t: t[..., :c].float().permute(0, 3, 1, 2)
    assert (y[..., c:] == 0).all()
    assert (nchw(y) - yr).abs().max().item() <= 2e-2 * yr.abs().max().item()
    assert (nchw(x.grad) - xr.grad).abs().max().item() <= 2e-2 * xr.grad.abs().max().item() + 1e-3
    assert (gam.grad - gr.grad).abs().max().item() <= 1e-2 * gr.grad.abs().max().item()
    assert (bet.grad - br.grad).abs().max().item() <= 1e-2 * br.grad.abs().max().item()
    if with_res:
        assert (nchw(res.grad) - rr.grad).abs().max().item() <= 1e-2 * rr.grad.abs().max().item() + 1e-3
    if training:
        assert torch.allclose(rm, rm2, rtol=1e-4, atol=1e-5) and torch.allclose(rv, rv2, rtol=1e-3, atol=1e-5)
