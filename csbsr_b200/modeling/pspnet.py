"""PSPNet (dilated ResNet-34 backbone) eval forward on the tcgen05 conv engine.

Mirrors PSPNet.forward (reference model/modeling/pspnet_pytorch/pspnet.py:95-123), PSPModule (:23-41),
PSPUpsample (:44-57) and the ResNet extractor (extractors.py:112-161, BasicBlock :41-70).  Eval-mode
BatchNorm is folded into the conv weights (scale) and bias (shift); ReLU / PReLU / sigmoid and the
residual add of BasicBlock are conv epilogues; torch.cat of the pyramid priors is replaced by writes
into channel slices of one 2560-channel buffer; Dropout2d is the identity in eval mode.
"""
import torch

from .. import kernels as K
from ..kernels import ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, Fmap
from .kbpn import _Workspace
from .params import RESNET34_LAYERS


class PSPNetEngine:
    def __init__(self, sizes=(1, 2, 3, 6), device="cuda"):
        self.sizes = sizes
        self.device = device
        self.ws = _Workspace(device)
        self.p = None

    def load(self, sd, prefix="segmentation_model."):
        dev = "cpu"                       # pack on the host; K.to_device ships the packed tensors (no device kernels at load time)
        g = lambda k: sd[prefix + k].detach().to(dev, torch.float32)

        def bn_fold(p, conv_bias=None, eps=1e-5):
            scale = g(p + ".weight") / torch.sqrt(g(p + ".running_var") + eps)
            shift = g(p + ".bias") - g(p + ".running_mean") * scale
            if conv_bias is not None:
                shift = shift + conv_bias * scale
            return scale, shift

        P = {}
        w = g("feats.conv1.weight")                                   # [64,3,7,7] on the patchified (147 -> 192) input
        sc, sh = bn_fold("feats.bn1")
        P["conv1"] = K.pack_conv(w.permute(0, 2, 3, 1).reshape(64, -1, 1, 1), sh, scale=sc)
        for li, (planes, blocks, stride, dil) in enumerate(RESNET34_LAYERS, 1):
            for b in range(blocks):
                bp = "feats.layer%d.%d" % (li, b)
                st = stride if b == 0 else 1
                d = 1 if b == 0 else dil
                sc, sh = bn_fold(bp + ".bn1")
                P[bp + ".c1"] = K.pack_conv(g(bp + ".conv1.weight"), sh, stride=st, padding=d, dilation=d, scale=sc)
                sc, sh = bn_fold(bp + ".bn2")
                P[bp + ".c2"] = K.pack_conv(g(bp + ".conv2.weight"), sh, padding=d, dilation=d, scale=sc)
                if (prefix + bp + ".downsample.0.weight") in sd:
                    sc, sh = bn_fold(bp + ".downsample.1")
                    P[bp + ".ds"] = K.pack_conv(g(bp + ".downsample.0.weight"), sh, stride=st, scale=sc)
        for i, _ in enumerate(self.sizes):
            P["psp%d" % i] = K.pack_conv(g("psp.stages.%d.1.weight" % i))
        P["bottleneck"] = K.pack_conv(g("psp.bottleneck.weight"), g("psp.bottleneck.bias"))
        for name in ("up_1", "up_2", "up_3"):
            sc, sh = bn_fold(name + ".conv.1", g(name + ".conv.0.bias"))
            P[name] = (K.pack_conv(g(name + ".conv.0.weight"), sh, padding=1, scale=sc),
                       float(sd[prefix + name + ".conv.2.weight"].detach().float().reshape(-1)[0]))
        # ---- BlurSkip (PSPNet_BlurSkip, pspnet.py:143-160; SFTLikeBlock blocks.py:105-120): the 441 conditioning channels
        # are spatially constant, so their half of conv 0 becomes a per-sample 3x3-border-class bias (as in KBPN's SFT)
        self.blur_layers = 0
        while (prefix + "blur_skip.%d.conv_scale.0.layer.weight" % (2 * self.blur_layers)) in sd:
            i = self.blur_layers
            for br in ("scale", "shift"):
                bp = "blur_skip.%d.conv_%s" % (2 * i, br)
                w0, b0 = g(bp + ".0.layer.weight"), g(bp + ".0.layer.bias")
                cc = w0.shape[0]
                cc_pad = K.round_up(cc, 64)
                P["bs%d.%s.0f" % (i, br)] = (K.pack_conv(w0[:, :64].contiguous(), padding=1, cout_pad=cc_pad),
                                             float(sd[prefix + bp + ".0.act.weight"].detach().float().reshape(-1)[0]))
                P["bs%d.%s.0k" % (i, br)] = K.pack_conv(w0[:, 64:].contiguous(), b0, padding=1, cout_pad=cc_pad,
                                                        cin_pad=K.round_up(cc - 64, 64))
                P["bs%d.%s.1" % (i, br)] = K.pack_conv(g(bp + ".1.layer.weight"), g(bp + ".1.layer.bias"), padding=1,
                                                       cin_pad=cc_pad)
            bp = "blur_skip.%d" % (2 * i + 1)
            sc, sh = bn_fold(bp + ".norm")
            P["bs%d.conv" % i] = K.pack_conv(g(bp + ".layer.weight"), sh, padding=1, scale=sc)
            self.blur_layers += 1
            self.blur_cc_pad, self.blur_cond = cc_pad, cc - 64
        P["final"] = K.pack_conv(g("final.0.weight"), g("final.0.bias"))
        sc, sh = bn_fold("aux.1")
        P["aux0"] = K.pack_conv(g("aux.0.weight"), sh, padding=1, scale=sc)
        P["aux4"] = K.pack_conv(g("aux.4.weight"), g("aux.4.bias"))
        self.p = K.to_device(P, self.device)
        return self

    def forward(self, img, mean=None, rstd=None, clamp01=False, kvec=None):
        """img: fp32 [B,3,H,W]; when mean/rstd ([B*3]) are given the input is clamp/instance-normalised on the
        fly while it is gathered for conv1 (build_model.py:135-146).  Returns (seg, aux) fp32 [B,1,H,W]."""
        P, ws = self.p, self.ws
        B, _, H, W = img.shape
        H2, W2 = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
        xp = K.patchify(img, ws.fmap("p192", B, H2, W2, 192), 7, 7, 2, 3, mean, rstd, clamp01)
        c1 = K.conv(xp, P["conv1"], ws.fmap("c1", B, H2, W2, 64), act=ACT_RELU)
        H4, W4 = (H2 + 2 - 3) // 2 + 1, (W2 + 2 - 3) // 2 + 1
        x = K.maxpool3s2(c1, ws.fmap("l1_a", B, H4, W4, 64))
        h, w = H4, W4
        x3 = None
        n_sizes = len(self.sizes)
        for li, (planes, blocks, stride, dil) in enumerate(RESNET34_LAYERS, 1):
            for b in range(blocks):
                bp = "feats.layer%d.%d" % (li, b)
                st = stride if b == 0 else 1
                oh, ow = (h - 1) // st + 1, (w - 1) // st + 1
                t = K.conv(x, P[bp + ".c1"], ws.fmap("l%d_t" % li, B, oh, ow, planes), act=ACT_RELU)
                res = x
                if (bp + ".ds") in P:
                    res = K.conv(x, P[bp + ".ds"], ws.fmap("l%d_ds" % li, B, oh, ow, planes))
                last = (li == 4 and b == blocks - 1)
                if last:      # final features go straight into the pyramid concat buffer (pspnet.py:40)
                    cat = ws.fmap("psp_cat", B, oh, ow, planes * (n_sizes + 1))
                    out = cat.window(planes * n_sizes, planes)
                else:
                    out = ws.fmap("l%d_%s" % (li, "b" if (b % 2 == 0) else "a"), B, oh, ow, planes)
                x = K.conv(t, P[bp + ".c2"], out, act=ACT_RELU, r0=res)
                h, w = oh, ow
            if li == 3:
                x3 = x
        feats, C = x, x.c
        # ---- pyramid pooling (pspnet.py:36-41)
        for i, s in enumerate(self.sizes):
            pooled = K.adaptive_avgpool(feats, ws.fmap("pool%d" % s, B, s, s, C), s)
            prior = K.conv(pooled, P["psp%d" % i], ws.fmap("prior%d" % s, B, s, s, C))
            K.bilinear(prior, cat.window(i * C, C))
        y = K.conv(cat, P["bottleneck"], ws.fmap("bott", B, h, w, 1024), act=ACT_RELU)
        # ---- three x2 upsample + conv + BN + PReLU blocks (pspnet.py:52-57, 106-113)
        for name, co in (("up_1", 256), ("up_2", 64), ("up_3", 64)):
            h, w = 2 * h, 2 * w
            up = K.bilinear(y, ws.fmap(name + "_in", B, h, w, y.c))
            y = K.conv(up, P[name][0], ws.fmap(name + "_out", B, h, w, co), act=ACT_LEAKY, slope=P[name][1])
        if self.blur_layers and kvec is not None:
            y = self._blur_skip(y, kvec, B, h, w)
        seg = torch.empty((B, 1, h, w), dtype=torch.float32, device=img.device)
        K.conv(y, P["final"], seg, act=ACT_SIGMOID)
        # ---- auxiliary head on layer3 features (pspnet.py:118-122)
        a = K.conv(x3, P["aux0"], ws.fmap("aux_t", B, x3.h, x3.w, 256), act=ACT_RELU)
        a1 = K.conv(a, P["aux4"], ws.f32("aux_small", B, 1, x3.h, x3.w), act=ACT_SIGMOID)
        aux = torch.empty((B, 1, H, W), dtype=torch.float32, device=img.device)
        K.bilinear_f32(a1, aux, align_corners=True)
        if (h, w) != (H, W):
            raise K._lib.CsbsrError("PSPNet output %dx%d != input %dx%d (input must be a multiple of 8)" % (h, w, H, W))
        return seg, aux

    def _blur_skip(self, p0, kvec, B, h, w):
        """p + BlurSkip(p, kernel) (pspnet.py:189-198): 2 x [SFTLikeBlock(64 + 441 -> 64), Conv3x3 + BN + ReLU]."""
        from ..kernels import F32Map
        P, ws = self.p, self.ws
        ccp, cond_pad = self.blur_cc_pad, K.round_up(self.blur_cond, 64)
        small = K.broadcast_vec(kvec, ws.fmap("bs_k3", B, 3, 3, cond_pad))
        t = p0
        for i in range(self.blur_layers):
            branch = {}
            for br in ("shift", "scale"):
                cb = K.conv(small, P["bs%d.%s.0k" % (i, br)], F32Map(ws.f32("bs_k3bias", B, 3, 3, ccp)))
                pc0, slope = P["bs%d.%s.0f" % (i, br)]
                u = K.conv(t, pc0, ws.fmap("bs_t", B, h, w, ccp), bias=cb.t, bias_sn=9 * ccp, bias_sc=ccp, cls_bw=1,
                           act=ACT_LEAKY, slope=slope)
                if br == "shift":
                    branch = K.conv(u, P["bs%d.shift.1" % i], ws.fmap("bs_shift", B, h, w, 64))
                else:
                    t2 = K.conv(u, P["bs%d.scale.1" % i], ws.fmap("bs_sft%d" % (i % 2), B, h, w, 64), act=ACT_SIGMOID, rm=t,
                                r1=branch)
            last = (i == self.blur_layers - 1)
            t = K.conv(t2, P["bs%d.conv" % i], ws.fmap("bs_out%d" % (i % 2), B, h, w, 64), act=ACT_RELU,
                       r1=p0 if last else None)
        return t
