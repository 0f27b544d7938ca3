"""TEST TOOL (CPU): how well conditioned is a training-step fixture?  Runs the fp32 oracle step twice on the fixture's inputs --
exact, and with every conv / transposed-conv operand rounded to bf16 (the ONLY deviation the tcgen05 engine makes by design) --
and prints the decorrelation between the two: if the fp32 oracle itself moves by X under operand rounding, a GPU-vs-reference
tolerance below X checks nothing but luck.

  python tests/tools/conditioning_probe.py tests/golden/train_step_b8.npz [--bn-eval] [--damp 0.1] [--save out.npz]

--save writes the gradient norms of BOTH oracle runs (exact fp32 / bf16-rounded operands): the GPU parity test accepts a tensor
whose norm is within tolerance of either -- the second run is the same pinned oracle with the engine's documented operand
precision, and it is what decides the ill-conditioned tensors (scalar PReLU slopes whose sum cancels, kb.sr_reconst).
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from csbsr_b200.modeling import params as P  # noqa: E402
from oracle import train_ref  # noqa: E402

DAMP = float(sys.argv[sys.argv.index("--damp") + 1]) if "--damp" in sys.argv else 1.0
NAMES = ["sr_model.feat.0.weight", "sr_model.predictor.feat_ext.2.layer.weight",
         "sr_model.back_projection_stages.0.up.up_conv1.layer.weight",
         "sr_model.back_projection_stages.1.sft.SFT_scale_conv0.weight",
         "sr_model.back_projection_stages.2.kb.kernel_predictor.fe_cat.2.layer.weight",
         "sr_model.back_projection_stages.3.kb.sr_reconst.layer.weight", "sr_model.output_conv.layer.weight",
         "segmentation_model.feats.conv1.weight", "segmentation_model.feats.layer3.2.conv1.weight",
         "segmentation_model.psp.bottleneck.weight", "segmentation_model.up_1.conv.0.weight",
         "segmentation_model.final.0.weight", "segmentation_model.aux.4.bias"]


def run(g, rounded, bn_train):
    sd = P.synth_state_dict(P.kbpn_param_shapes(), prefix="sr_model.")
    sd.update(P.synth_state_dict(P.pspnet_param_shapes(), prefix="segmentation_model."))
    if DAMP != 1.0:
        for k in sd:
            if ".feats.layer" in k and k.endswith("bn2.weight"):
                sd[k] = sd[k] * DAMP
    for k, v in sd.items():
        if v.is_floating_point():
            v.requires_grad_(True)
    lr, hr, mask, kgt = (torch.from_numpy(g[k]) for k in ("lr", "hr", "mask", "kgt"))
    c2, ct2 = F.conv2d, F.conv_transpose2d
    if rounded:
        r = lambda t: t.to(torch.bfloat16).to(torch.float32) + (t - t.detach()) * 0 if False else _ste(t)
        F.conv2d = lambda x, w, *a, **k: c2(r(x), r(w), *a, **k)
        F.conv_transpose2d = lambda x, w, *a, **k: ct2(r(x), r(w), *a, **k)
    try:
        loss, seg_loss, sr_loss, sr, seg, aux = train_ref.train_forward(sd, lr, hr, mask, kgt, float(g["alpha"]), bn_train=bn_train)
        loss.backward()
    finally:
        F.conv2d, F.conv_transpose2d = c2, ct2
    return loss.item(), seg.detach(), sr.detach(), {k: sd[k].grad.clone() for k in sd if sd[k].grad is not None}


class _STE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, gy):
        return gy.to(torch.bfloat16).to(torch.float32)      # dgrad / wgrad operands are bf16 too


def _ste(t):
    return _STE.apply(t)


if __name__ == "__main__":
    g = np.load(sys.argv[1])
    bn_train = "--bn-eval" not in sys.argv
    torch.set_num_threads(os.cpu_count())
    l0, seg0, sr0, g0 = run(g, False, bn_train)
    l1, seg1, sr1, g1 = run(g, True, bn_train)
    print("loss exact %.6f  bf16-operand %.6f  (fixture %.6f)" % (l0, l1, float(g["loss"])))
    print("seg mean abs diff %.5f  max %.5f ; sr max diff %.5f" % ((seg0 - seg1).abs().mean(), (seg0 - seg1).abs().max(), (sr0 - sr1).abs().max()))
    if "--save" in sys.argv:
        names = sorted(g0)
        np.savez_compressed(sys.argv[sys.argv.index("--save") + 1], names=np.array(names),
                            norm_exact=np.array([float(g0[k].double().norm()) for k in names]),
                            norm_bf16ops=np.array([float(g1[k].double().norm()) for k in names]),
                            loss_exact=np.float64(l0), loss_bf16ops=np.float64(l1), damp=np.float32(DAMP))
    gmax = max(float(v.double().norm()) for v in g0.values())
    for k in g0:
        a, b = g0[k].flatten().double(), g1[k].flatten().double()
        ratio = float(b.norm() / (a.norm() + 1e-30))
        if k in NAMES or abs(ratio - 1) > 0.15:
            print("  %-80s cos %.5f  norm ratio %.3f  |g| %.3e (max %.3e)%s" % (k, float(a @ b / (a.norm() * b.norm() + 1e-30)), ratio,
                  float(a.norm()), gmax, "" if k in NAMES else "   <-- outside 15 %"))
