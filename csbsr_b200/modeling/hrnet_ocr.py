"""HRNet-W48 + OCR segmentation head, eval forward on the tcgen05 conv engine (config #4 of BASELINE.json).

Mirrors HRNet_W48_OCR.forward (reference model/modeling/hrnet_ocr/nets/hrnet.py:137-158), HighResolutionNet.forward
(backbones/hrnet/hrnet_backbone.py:514-572), HighResolutionModule.forward (:265-290), SpatialGather_Module
(modules/spatial_ocr_block.py:49-66), _ObjectAttentionBlock.forward (:172-196) and SpatialOCR_Module.forward (:281-303).

B200-first restructurings (exact in real arithmetic):
  * eval BatchNorm folded into conv weights / bias; ReLU and residual adds are conv epilogues;
  * multi-resolution fusion `y_i = relu(sum_j f_ij(x_j))` accumulates in place: up-sampled terms through
    csbsr_bilinear_add_nhwc, strided-conv terms through the conv epilogue's pre-activation residual;
  * with one object class (K = 1) the pixel-object attention map is softmax over a single element, i.e. exactly 1, so
    f_pixel / f_object do not influence the output; the object context is one vector per image that enters the final
    1x1 conv (cat([context, feats]) -> 512) as a per-sample bias.
Channels 48 / 96 / 720 are zero-padded to 64 / 128 / 768 (TMA boxes and UMMA K slabs are 64 channels wide)."""
import torch

from .. import kernels as K
from ..kernels import ACT_NONE, ACT_RELU, F32Map, Fmap
from .kbpn import _Workspace
from .params import HRNET48_STAGES


def _cp(c):
    return K.round_up(c, 64)


class HRNetOCREngine:
    def __init__(self, device="cuda"):
        self.device = device
        self.ws = _Workspace(device)
        self.p = None

    # ------------------------------------------------------------------ weights
    def load(self, sd, prefix="segmentation_model."):
        dev = "cpu"                       # pack on the host; K.to_device ships the packed tensors (no device kernels at load time)
        g = lambda k: sd[prefix + k].detach().to(dev, torch.float32)

        def fold(conv, bn, stride=1, padding=0, cin_pad=None):
            scale = g(bn + ".weight") / torch.sqrt(g(bn + ".running_var") + 1e-5)
            shift = g(bn + ".bias") - g(bn + ".running_mean") * scale
            if (prefix + conv + ".bias") in sd:
                shift = shift + g(conv + ".bias") * scale
            w = g(conv + ".weight")
            return K.pack_conv(w, shift, stride=stride, padding=padding, scale=scale, cout_pad=_cp(w.shape[0]), cin_pad=cin_pad)

        P = {}
        b = "backbone."
        w1 = g(b + "conv1.weight")                                           # 3x3 s2 on the patchified (27 -> 64) input
        sc = g(b + "bn1.weight") / torch.sqrt(g(b + "bn1.running_var") + 1e-5)
        sh = g(b + "bn1.bias") - g(b + "bn1.running_mean") * sc
        P["conv1"] = K.pack_conv(w1.permute(0, 2, 3, 1).reshape(64, -1, 1, 1), sh, scale=sc)
        P["conv2"] = fold(b + "conv2", b + "bn2", stride=2, padding=1)
        for i in range(4):
            bp = b + "layer1.%d" % i
            P["l1.%d.c1" % i] = fold(bp + ".conv1", bp + ".bn1")
            P["l1.%d.c2" % i] = fold(bp + ".conv2", bp + ".bn2", padding=1)
            P["l1.%d.c3" % i] = fold(bp + ".conv3", bp + ".bn3")
        P["l1.ds"] = fold(b + "layer1.0.downsample.0", b + "layer1.0.downsample.1")
        pre = (256,)
        for si, (modules, chans) in enumerate(HRNET48_STAGES, 2):
            t = b + "transition%d" % (si - 1)
            for i, c in enumerate(chans):
                if i < len(pre):
                    if c != pre[i]:
                        P["t%d.%d" % (si, i)] = fold(t + ".%d.0" % i, t + ".%d.1" % i, padding=1)
                else:
                    P["t%d.%d" % (si, i)] = fold(t + ".%d.0.0" % i, t + ".%d.0.1" % i, stride=2, padding=1)
            for m in range(modules):
                mp = b + "stage%d.%d" % (si, m)
                key = "s%d.%d" % (si, m)
                for bi in range(len(chans)):
                    for k in range(4):
                        bp = mp + ".branches.%d.%d" % (bi, k)
                        P[key + ".b%d.%d.c1" % (bi, k)] = fold(bp + ".conv1", bp + ".bn1", padding=1)
                        P[key + ".b%d.%d.c2" % (bi, k)] = fold(bp + ".conv2", bp + ".bn2", padding=1)
                for i in range(len(chans)):
                    for j in range(len(chans)):
                        fp = mp + ".fuse_layers.%d.%d" % (i, j)
                        if j > i:
                            P[key + ".f%d.%d" % (i, j)] = fold(fp + ".0", fp + ".1")
                        elif j < i:
                            for k in range(i - j):
                                P[key + ".f%d.%d.%d" % (i, j, k)] = fold(fp + ".%d.0" % k, fp + ".%d.1" % k, stride=2, padding=1)
            pre = chans
        P["aux0"] = fold("aux_head.0", "aux_head.1.0", padding=1, cin_pad=768)
        P["aux2"] = K.pack_conv(g("aux_head.2.weight"), g("aux_head.2.bias"), cin_pad=768)
        P["conv3x3"] = fold("conv3x3.0", "conv3x3.1.0", padding=1, cin_pad=768)
        o = "ocr_distri_head.object_context_block."
        P["f_down"] = fold(o + "f_down.0", o + "f_down.1.0")
        P["f_up"] = fold(o + "f_up.0", o + "f_up.1.0")
        # conv_bn_dropout over cat([context, feats]): context half -> per-sample bias, feats half -> the big 1x1 conv
        q = "ocr_distri_head.conv_bn_dropout."
        scale = g(q + "1.0.weight") / torch.sqrt(g(q + "1.0.running_var") + 1e-5)
        shift = g(q + "1.0.bias") - g(q + "1.0.running_mean") * scale + g(q + "0.bias") * scale
        wq = g(q + "0.weight")
        P["ocr_ctx"] = K.pack_conv(wq[:, :512].contiguous(), shift, scale=scale)
        P["ocr_feat"] = K.pack_conv(wq[:, 512:].contiguous(), None, scale=scale)
        P["cls"] = K.pack_conv(g("cls_head.weight"), g("cls_head.bias"))
        self.p = K.to_device(P, self.device)
        return self

    # ------------------------------------------------------------------ forward
    def _basic_blocks(self, key, bi, x, tag):
        """4 BasicBlocks (hrnet_backbone.py:33-66) on one branch; returns the output Fmap."""
        ws, P = self.ws, self.p
        n, h, w, c = x.n, x.h, x.w, x.pitch
        for k in range(4):
            t = K.conv(x, P[key + ".b%d.%d.c1" % (bi, k)], ws.fmap(tag + "_t", n, h, w, c), act=ACT_RELU)
            out = ws.fmap(tag + ("_a" if k % 2 == 0 else "_b"), n, h, w, c)
            x = K.conv(t, P[key + ".b%d.%d.c2" % (bi, k)], out, act=ACT_RELU, r0=x)
        return x

    def forward(self, img, mean=None, rstd=None, clamp01=False, kvec=None):
        """img fp32 [B,3,H,W] (optionally clamp + instance-normalised on the fly) -> (seg, aux) fp32 [B,1,H,W]."""
        P, ws = self.p, self.ws
        B, _, H, W = img.shape
        H2, W2 = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        H4, W4 = (H2 - 1) // 2 + 1, (W2 - 1) // 2 + 1
        xp = K.patchify(img, ws.fmap("h_p64", B, H2, W2, 64), 3, 3, 2, 1, mean, rstd, clamp01)
        x = K.conv(xp, P["conv1"], ws.fmap("h_c1", B, H2, W2, 64), act=ACT_RELU)
        x = K.conv(x, P["conv2"], ws.fmap("h_c2", B, H4, W4, 64), act=ACT_RELU)
        # ---- layer1: 4 Bottlenecks 64 -> 256 (hrnet_backbone.py:69-105)
        for i in range(4):
            t = K.conv(x, P["l1.%d.c1" % i], ws.fmap("h_l1t1", B, H4, W4, 64), act=ACT_RELU)
            t = K.conv(t, P["l1.%d.c2" % i], ws.fmap("h_l1t2", B, H4, W4, 64), act=ACT_RELU)
            res = K.conv(x, P["l1.ds"], ws.fmap("h_l1ds", B, H4, W4, 256)) if i == 0 else x
            x = K.conv(t, P["l1.%d.c3" % i], ws.fmap("h_l1" + ("a" if i % 2 == 0 else "b"), B, H4, W4, 256), act=ACT_RELU, r0=res)
        ys, pre = [x], (256,)
        for si, (modules, chans) in enumerate(HRNET48_STAGES, 2):
            sizes = [((H4 - 1) // (1 << i) + 1, (W4 - 1) // (1 << i) + 1) for i in range(len(chans))]
            xs = []
            for i, c in enumerate(chans):                            # transition layers (:402-447)
                key = "t%d.%d" % (si, i)
                if key in P:
                    src = ys[i] if i < len(pre) else ys[-1]
                    xs.append(K.conv(src, P[key], ws.fmap("h_tr%d_%d" % (si, i), B, sizes[i][0], sizes[i][1], _cp(c)), act=ACT_RELU))
                else:
                    xs.append(ys[i])
            for m in range(modules):
                key = "s%d.%d" % (si, m)
                xs = [self._basic_blocks(key, bi, xs[bi], "h_s%d_b%d" % (si, bi)) for bi in range(len(chans))]
                fused = []
                nb = len(chans)
                for i in range(nb):                                  # HighResolutionModule.forward fusion (:274-288)
                    out = ws.fmap("h_s%d_f%d_%d" % (si, i, m % 2), B, sizes[i][0], sizes[i][1], _cp(chans[i]))
                    acc = xs[i]
                    terms = [("up", j) for j in range(i + 1, nb)] + [("down", j) for j in range(i)]
                    for ti, (kind, j) in enumerate(terms):
                        last = ti == len(terms) - 1
                        if kind == "up":
                            tmp = K.conv(xs[j], P[key + ".f%d.%d" % (i, j)],
                                         ws.fmap("h_fu%d_%d" % (j, _cp(chans[i])), B, sizes[j][0], sizes[j][1], _cp(chans[i])))
                            acc = K.bilinear_add(tmp, acc, out, align_corners=True, relu=last)
                        else:
                            t = xs[j]
                            for k in range(i - j):
                                pc = P[key + ".f%d.%d.%d" % (i, j, k)]
                                if k == i - j - 1:
                                    acc = K.conv(t, pc, out, act=ACT_RELU if last else ACT_NONE, r0=acc)
                                else:
                                    t = K.conv(t, pc, ws.fmap("h_fd%d_%d" % (j + k + 1, _cp(chans[j])), B, sizes[j + k + 1][0],
                                                              sizes[j + k + 1][1], _cp(chans[j])), act=ACT_RELU)
                    fused.append(acc)
                xs = fused
            ys, pre = xs, chans
        # ---- OCR head (hrnet.py:137-158)
        h, w = ys[0].h, ys[0].w
        cat = ws.fmap("h_cat", B, h, w, 768)
        off = 0
        for y, c in zip(ys, HRNET48_STAGES[-1][1]):
            K.bilinear(y.window(0, c), cat.window(off, c), align_corners=True)
            off += c
        a = K.conv(cat, P["aux0"], ws.fmap("h_aux", B, h, w, 768), act=ACT_RELU)
        aux_logit = K.conv(a, P["aux2"], ws.f32("h_auxl", B, 1, h, w))
        f = K.conv(cat, P["conv3x3"], ws.fmap("h_feat", B, h, w, 512), act=ACT_RELU)
        ctx = K.softmax_gather(aux_logit, f, ws.f32("h_ctx", B, 512))
        # object context: softmax over the single class is 1 -> context = f_up(f_down(proxy)), one vector per image
        cv = K.nchw_to_nhwc(ctx.view(B, 512, 1, 1), ws.fmap("h_ctxv", B, 1, 1, 512))
        cv = K.conv(cv, P["f_down"], ws.fmap("h_ctxd", B, 1, 1, 256), act=ACT_RELU)
        cv = K.conv(cv, P["f_up"], ws.fmap("h_ctxu", B, 1, 1, 512), act=ACT_RELU)
        cb = K.conv(cv, P["ocr_ctx"], F32Map(ws.f32("h_ctxb", B, 1, 1, 512)))
        f2 = K.conv(f, P["ocr_feat"], ws.fmap("h_feat2", B, h, w, 512), bias=cb.t, bias_sn=512, bias_sc=0, cls_bw=0, act=ACT_RELU)
        logit = K.conv(f2, P["cls"], ws.f32("h_logit", B, 1, h, w))
        seg = torch.empty((B, 1, H, W), dtype=torch.float32, device=img.device)
        aux = torch.empty((B, 1, H, W), dtype=torch.float32, device=img.device)
        K.bilinear_f32_sigmoid(logit, seg, align_corners=True)
        K.bilinear_f32_sigmoid(aux_logit, aux, align_corners=True)
        return seg, aux
