"""smoke(): one small invocation of the whole hot path on cuda:0, checked against the oracle (test infrastructure)."""
import os
import sys

import numpy as np
import torch


def run():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from csbsr_b200 import _lib
    from csbsr_b200.config import cfg
    from csbsr_b200.data import degrade as G
    from csbsr_b200.engine import inference as E
    from csbsr_b200.modeling.build_model import JointModel
    from csbsr_b200.utils import synth
    from oracle import degrade_ref, metrics_ref, torch_ref

    torch.cuda.set_device(0)
    _lib.lib()
    c = cfg.clone()
    c.merge_from_file(os.path.join(root, "config", "config_csbsr_pspnet.yaml"))
    model = JointModel(c)
    sd = synth.model_state_dict()
    model.load_state_dict(sd, strict=True)
    hr, mask = synth.batch(0, 2, 96)
    params = synth.degradation_params(2)
    lr, kern = G.degrade(hr, params)
    sr, seg, kp = model(lr, torch.zeros(2, 1, 7, 7))
    r = E.seg_metrics(seg, mask, with_hd=True)
    torch.cuda.synchronize()
    n_eval = _lib.LAUNCHES

    # ---- roll call of the training kernels right after the eval path (conv forward, dgrad on the flipped packing, wgrad,
    # weight packing), so that a bounded launch trace of smoke() reaches them: one conv Function forward + backward,
    # checked against torch autograd further down
    from csbsr_b200 import autograd as A
    gen = torch.Generator().manual_seed(5)
    cx0 = torch.randn(2, 64, 20, 24, generator=gen).to(torch.bfloat16).float().cuda()
    cw0 = (torch.randn(128, 64, 3, 3, generator=gen) * 0.06).to(torch.bfloat16).float().cuda()
    cup = torch.randn(2, 128, 20, 24, generator=gen).to(torch.bfloat16).float().cuda()
    cx, cw = cx0.clone().requires_grad_(True), cw0.clone().requires_grad_(True)
    cy = A.conv2d(A.to_nhwc(cx), cw, None, stride=1, padding=1)
    cy.backward(A.to_nhwc(cup))
    torch.cuda.synchronize()

    lr_ref, k_ref, _ = degrade_ref.degrade(hr, params)
    assert (lr.cpu() - lr_ref).abs().max().item() <= 2e-6, "degrade mismatch"
    # the oracle runs on the HOST (true fp32, no TF32, and no library kernels in a launch trace of smoke())
    with torch.no_grad():
        sr_ref, seg_ref, kp_ref, _ = torch_ref.joint_forward(sd, lr.cpu())
    e_sr, e_seg = (sr.cpu() - sr_ref).abs().max().item(), (seg.cpu() - seg_ref).abs().max().item()
    assert e_sr <= 3e-2 and e_seg <= 2e-2, "network mismatch sr %g seg %g" % (e_sr, e_seg)
    inter, union = metrics_ref.iou_counts(seg.cpu().numpy(), mask.numpy())
    hd, msd = metrics_ref.distance_metrics(seg.cpu().numpy(), mask.numpy(), 50)
    assert np.array_equal(r["inter"], inter) and np.array_equal(r["union"], union), "AIU counts mismatch"
    assert np.array_equal(r["hd"], hd) and np.array_equal(r["msd"], msd), "HD/MSD mismatch"
    cxr, cwr = cx0.cpu().requires_grad_(True), cw0.cpu().requires_grad_(True)
    torch.nn.functional.conv2d(cxr, cwr, None, padding=1).backward(cup.cpu())
    rel = lambda a, b: (a.cpu() - b).abs().max().item() / (b.abs().max().item() + 1e-12)
    assert rel(cx.grad, cxr.grad) <= 1e-2 and rel(cw.grad, cwr.grad) <= 1e-2, "conv dgrad / wgrad mismatch"

    # ---- one tiny joint training step (forward, loss, backward through the conv dgrad / wgrad kernels, fused Adam),
    # loss checked against the fp32 oracle's (BatchNorm on running statistics and Dropout2d off so both are deterministic)
    from csbsr_b200.engine.losses import calc_loss
    from csbsr_b200.engine.optim import FusedAdam
    from csbsr_b200.modeling.build_model import JointModelWithLoss
    from oracle import train_ref
    tc = c.clone()
    tc.SOLVER.SEG_FAIL_ORIENTED_WEIGHT4SS_AMP = 1.0
    tm = JointModelWithLoss(tc, num_train_ds=100, resume_iter=40000)
    tm.load_state_dict(sd, strict=True)
    tm.cuda().train()
    tm.dropout, tm.freeze_bn = False, True
    opt = FusedAdam(tm.parameters(), lr=tc.SOLVER.LR)
    hr_t, mask_t = synth.batch(4, 2, 64)
    lr_t, kern_t = G.degrade(hr_t, synth.degradation_params(2, seed=9))
    seg_loss, sr_loss, *_ = tm(40001, lr_t, sr_targets=hr_t, segment_targets=mask_t, kernel_targets=kern_t.unsqueeze(1))
    loss = calc_loss(sr_loss, seg_loss.mean(), tc.SOLVER.TASK_LOSS_WEIGHT, 40001, tc)
    loss.backward()
    gnorm = opt.flat_g.norm().item()
    opt.step()
    torch.cuda.synchronize()
    with torch.no_grad():
        ref_loss = train_ref.train_forward(sd, lr_t.cpu(), hr_t.cpu(), mask_t.cpu(), kern_t.cpu().unsqueeze(1), tm.ss_loss_fn.alpha,
                                           beta=tc.SOLVER.TASK_LOSS_WEIGHT, wf_amp=1.0, bn_train=False)[0].item()
    assert np.isfinite(gnorm) and gnorm > 0, "training step produced no gradient"
    assert abs(loss.item() - ref_loss) <= 3e-2 * abs(ref_loss), "train loss %g vs oracle %g" % (loss.item(), ref_loss)
    print("smoke ok: sr max-abs %.4f, seg max-abs %.4f, AIU %.4f, AHD(p50) %.3f, %d kernel launches (eval); "
          "train step loss %.5f (oracle %.5f), |grad| %.3e, %d launches"
          % (e_sr, e_seg, float(np.mean(r["iou"])), float(np.mean(r["hd"])), n_eval, loss.item(), ref_loss, gnorm,
             _lib.LAUNCHES - n_eval))
