"""Training-mode forward of KBPN + PSPNet as an autograd graph over the tcgen05 conv Functions.

Every dense conv / transposed conv runs forward, dgrad and wgrad on the csbsr_b200 engine (csbsr_b200/autograd.py), and
the glue between them -- activations, residual adds, SFT combine, concat, bilinear / adaptive / max pooling, Dropout2d,
BatchNorm, instance norm, layout changes -- runs forward and backward on the kernels of csrc/glue.cu, csrc/support.cu and
csrc/train.cu (csbsr_b200/glue.py); torch.autograd is the tape.  What is left to aten are O(B x 441)-sized vector ops on the
blur-kernel estimates and fp32 adds on 3-channel images.  Activations are NHWC bf16 with channels zero-padded to a multiple of
64; the SR image, the blur-kernel vectors and everything feeding the losses stay fp32.

Mirrors, in train mode and outside the pre-training phases (iteration >= SR_PRETRAIN_ITER[1]):
  KBPN.forward                      model/modeling/kbpn.py:84-116
  KernelBackProjectionStageWithSFT  kbpn.py:172-189, KBlock :382-412, UpBlock :464-469, DownBlock :484-489,
  SFTlayer :511-518, KernelPredictorLikeIKC :562-578, predictor_withGAP :320-341
  PSPNet.forward                    model/modeling/pspnet_pytorch/pspnet.py:95-123 (BatchNorm batch statistics,
                                    Dropout2d 0.3 / 0.15 / 0.1), ResNet extractors.py:150-161
"""
import torch
import torch.nn.functional as F

from .. import glue as G
from ..autograd import batch_norm, conv2d, conv3x3_few_outputs, cpad, deconv8s4, prelu
from ..glue import to_nchw, to_nhwc
from .params import RESNET34_LAYERS

_EPI_ACT = {"relu": "relu", "lrelu": ("lrelu", 0.01)}          # activations that ride in the conv epilogue


def _cat(parts, real=None):
    """Concatenate NHWC maps along channels (the first real[i] channels of part i), zero-padded to a multiple of 64."""
    return G.concat(parts, real)


def _convblock(P, p, x, stride=1, padding=0, act=None):
    w_ = P[p + ".layer.weight"]
    if (act is None and stride == 1 and padding == 1 and tuple(w_.shape[2:]) == (3, 3) and w_.shape[0] <= 4 and w_.shape[1] >= 64
            and (p + ".layer.bias") not in P):
        return conv3x3_few_outputs(x, w_)                 # sr_reconst / output_conv: tap-expanded 1x1 GEMMs
    y = conv2d(x, P[p + ".layer.weight"], P.get(p + ".layer.bias"), stride=stride, padding=padding, act=_EPI_ACT.get(act))
    return prelu(y, P[p + ".act.weight"]) if act == "prelu" else y


def _deconvblock(P, p, x, act="prelu"):
    y = deconv8s4(x, P[p + ".layer.weight"], P.get(p + ".layer.bias"))
    return prelu(y, P[p + ".act.weight"]) if act == "prelu" else y


def _gap(x, c):
    return G.gap(x, c)                                                # [B, c] fp32


def _upscale_kernel(vec, k_out):
    k = int(round(vec.shape[1] ** 0.5))
    return F.interpolate(vec.view(vec.shape[0], 1, k, k), size=(k_out, k_out), mode="bicubic")


def _expand_vec(kvec, h, w):
    """(B, 441) fp32 -> NHWC bf16 [B, h, w, 448] conditioning map (differentiable)."""
    return _BroadcastVecFn.apply(kvec, h, w)


class _BroadcastVecFn(torch.autograd.Function):
    """fp32 [B, c] -> NHWC bf16 [B, h, w, cpad(c)] with out[b, y, x, :c] = v[b]; backward = per-sample sum over the pixels."""

    @staticmethod
    def forward(ctx, v, h, w):
        from .. import kernels as K
        b, c = v.shape
        out = torch.empty((b, h, w, cpad(c)), dtype=torch.bfloat16, device=v.device)
        K.broadcast_vec(v.contiguous().float(), K.Fmap(out))
        ctx.cfg = (c, h, w)
        return out

    @staticmethod
    def backward(ctx, dy):
        from .. import kernels as K
        c, h, w = ctx.cfg
        dy = dy.contiguous()
        g = torch.empty((dy.shape[0], c), dtype=torch.float32, device=dy.device)
        K.gap(K.Fmap(dy), g, c)
        return g * float(h * w), None, None


def _expand_classes(small, h, w, bw):
    """[B, 2bw+1, 2bw+1, C] responses per border class -> [B, h, w, C] (class of a position: 0..bw-1 at the start of an axis,
    bw in the interior, bw+1..2bw at the end); backward sums the gradient over the pixels of every class."""
    return G.expand_classes(small, h, w, bw)


def _kernel_predictor(P, p, sr_t, kvec, k_out):
    n, H, W, _ = sr_t.shape
    fsr = _convblock(P, p + ".fe_SR.0", sr_t, padding=1, act="relu")
    fsr = _convblock(P, p + ".fe_SR.1", fsr, act="lrelu")
    fsr = _convblock(P, p + ".fe_SR.2", fsr, padding=1, act="lrelu")
    fsr = _convblock(P, p + ".fe_SR.3", fsr, padding=1, act="lrelu")
    fsr = _convblock(P, p + ".fe_SR.4", fsr, padding=1, act="lrelu")
    # fe_kernel sees a spatially constant map (kbpn.py:572-573): two stacked zero-padded 3x3 convs respond identically
    # everywhere except within 2 px of the border -> evaluate them on a 5x5 image of border classes and gather
    # (exactly the reference's values; gradients reach the kernel vector and both convs through the gather)
    fh = _expand_vec(kvec, 5, 5)
    fh = _convblock(P, p + ".fe_kernel.0", fh, padding=1, act="lrelu")
    fh = _convblock(P, p + ".fe_kernel.1", fh, padding=1, act="lrelu")
    fh = _expand_classes(fh, H, W, 2)
    c = P[p + ".fe_SR.4.layer.weight"].shape[0]
    # cat(fsr[:c], fh[:c]) with c = 49: the padded maps are concatenated whole (64 + 64 channels) and the input channels of the
    # 1x1 fe_cat.0 weight are scattered to the padded positions instead (differentiable, a [32, 98] -> [32, 128] copy)
    cp_ = fsr.shape[3]
    d = _cat((fsr, fh))
    w_cat = P[p + ".fe_cat.0.layer.weight"]
    w_pad = torch.zeros((w_cat.shape[0], 2 * cp_, 1, 1), dtype=w_cat.dtype, device=w_cat.device)
    w_pad = torch.cat((w_cat[:, :c], w_pad[:, c:cp_], w_cat[:, c:], w_pad[:, cp_ + c:]), dim=1)
    d = conv2d(d, w_pad, None, act=_EPI_ACT["lrelu"])
    d = _convblock(P, p + ".fe_cat.1", d, padding=1, act="lrelu")
    d = _convblock(P, p + ".fe_cat.2", d, padding=1, act=None)
    delta = _upscale_kernel(_gap(d, c), k_out).reshape(n, k_out * k_out)
    return kvec + delta


def blur_per_sample(img, kvec, k_out, stride):
    """Depthwise cross-correlation of every sample with its own kernel (kbpn.py:395-402; sr_loss_functions.py:73-102).
    img fp32 [B,3,H,W], kvec [B, k*k] -> [B,3,H/stride,W/stride]; forward and both gradients on csbsr kernels."""
    from ..autograd import blur_per_sample as _blur
    return _blur(img, kvec, k_out, stride)


def _up_block(P, p, x):
    x = _convblock(P, p + ".conv", x, act="prelu")
    h0 = _deconvblock(P, p + ".up_conv1", x)
    l0 = _convblock(P, p + ".up_conv2", h0, stride=4, padding=2, act="prelu")
    h1 = _deconvblock(P, p + ".up_conv3", G.sub(l0, x))
    return G.add(h1, h0)


def _down_block(P, p, x):
    x = _convblock(P, p + ".conv", x, act="prelu")
    l0 = _convblock(P, p + ".down_conv1", x, stride=4, padding=2, act="prelu")
    h0 = _deconvblock(P, p + ".down_conv2", l0)
    l1 = _convblock(P, p + ".down_conv3", G.sub(h0, x), stride=4, padding=2, act="prelu")
    return G.add(l1, l0)


def _k_block(P, p, concat_h, h, x_lr, kvec, k_out, scale, predict_kernel=True):
    sr_t = _convblock(P, p + ".sr_reconst", concat_h, padding=1)
    # during the SR-module pre-training the (ground-truth) kernel passes through unchanged (kbpn.py:386-388)
    d_kernel = _kernel_predictor(P, p + ".kernel_predictor", sr_t, kvec, k_out) if predict_kernel else kvec
    vec = d_kernel / d_kernel.sum(dim=1, keepdim=True)
    pseudo_lr = blur_per_sample(to_nchw(sr_t, 3), vec, k_out, scale)
    e_h = _deconvblock(P, p + ".up_conv1", to_nhwc(pseudo_lr - x_lr))
    return G.add(h, e_h), vec


def _sft(P, p, feats, kvec):
    """SFTlayer.forward (kbpn.py:511-518).  conv0 acts on cat(features, kernel map); the 441 kernel-map channels are
    spatially constant, so their part of the (linear) conv is evaluated on a 3x3 image of border classes and added as a
    per-sample, per-class bias: conv0(cat(f, k)) = conv(f, W[:, :fc]) + gather(conv(k_3x3, W[:, fc:]) + b)."""
    n, h, w, fc = feats.shape
    cond = _expand_vec(kvec, 3, 3)

    def branch(name):
        w0 = P[p + ".SFT_%s_conv0.weight" % name]
        fcr = w0.shape[1] - kvec.shape[1]                 # real feature channels (fc is their padded count)
        t = conv2d(feats, w0, None, padding=1, cin_range=(0, fcr))
        tb = conv2d(cond, w0, P[p + ".SFT_%s_conv0.bias" % name], padding=1, cin_range=(fcr, kvec.shape[1]))
        t = G.leaky_relu(G.add(t, _expand_classes(tb, h, w, 1)), 0.1)
        return conv2d(t, P[p + ".SFT_%s_conv1.weight" % name], P[p + ".SFT_%s_conv1.bias" % name], padding=1)
    return G.sft_combine(feats, branch("scale"), branch("shift"))


def kbpn_forward(P, x_lr, num_stages=4, k_out=21, scale=4, prefix="sr_model.", gt_kernel=None):
    """x_lr fp32 [B,3,h,w] -> (sr fp32 [B,3,4h,4w], kernel vector fp32 [B, 441] (normalised, as KBlock returns it)).
    `gt_kernel` (B,1,k,k): SR-module pre-training (kbpn.py:89-91, 386-388) -- the ground-truth kernel replaces the initial
    prediction and the per-stage kernel predictors are skipped."""
    p = prefix
    f = to_nhwc(x_lr)
    for i in (0, 2, 4, 6):
        f = conv2d(f, P[p + "feat.%d.weight" % i], P[p + "feat.%d.bias" % i], padding=1, act="relu")
    init_f = f
    if gt_kernel is not None:
        kvec = gt_kernel.reshape(x_lr.shape[0], k_out * k_out).float()
    else:
        z = init_f
        for i in range(3):
            z = _convblock(P, p + "predictor.feat_ext.%d" % i, z, padding=1, act="prelu")
        ke2 = P[p + "predictor.feat_ext.2.layer.weight"].shape[0]
        ker = _upscale_kernel(_gap(z, ke2), k_out)
        kvec = (ker / ker.sum(dim=(2, 3), keepdim=True)).reshape(x_lr.shape[0], k_out * k_out)
    low, concat_h, concat_l = init_f, None, None
    for s in range(num_stages):
        sp = p + "back_projection_stages.%d" % s
        h = _up_block(P, sp + ".up", low)
        pre = h if concat_h is None else _cat((concat_h, h))
        h, kvec = _k_block(P, sp + ".kb", pre, h, x_lr, kvec, k_out, scale, predict_kernel=gt_kernel is None)
        concat_h = h if concat_h is None else _cat((concat_h, h))
        if s < num_stages - 1:
            low = _down_block(P, sp + ".down", concat_h)
            concat_l = low if concat_l is None else _cat((concat_l, low))
            low = _sft(P, sp + ".sft", concat_l, kvec)
    sr = to_nchw(_convblock(P, p + "output_conv", concat_h, padding=1), 3)
    from .. import kernels as K
    up = torch.empty_like(sr)
    K.bicubic_upsample(x_lr.contiguous(), up, scale)                 # nn.Upsample(bicubic) of the (constant) LR input, kbpn.py:70,113
    return sr + up, kvec


# ---------------------------------------------------------------------------------------------- PSPNet (train mode)
def _bn(P, p, x, training, momentum=0.1, relu=False, res=None):
    """BatchNorm2d (+ fused residual add and ReLU) on an NHWC tensor with the csbsr_bn_* kernels; train mode uses batch
    statistics and updates the running buffers in place.  Channel-padded tensors stay zero in the padding channels."""
    y = batch_norm(x, P[p + ".weight"], P[p + ".bias"], P[p + ".running_mean"], P[p + ".running_var"], training, momentum,
                   1e-5, relu=relu, res=res)
    if training and (p + ".num_batches_tracked") in P:
        _TRACKED.append(P[p + ".num_batches_tracked"])
    return y


_TRACKED = []          # num_batches_tracked buffers touched by the current forward: advanced by one batched launch at its end


def _flush_tracked():
    if _TRACKED:
        torch._foreach_add_(_TRACKED, 1)
        _TRACKED.clear()


def _drop(x, p, state, c=None):
    """nn.Dropout2d(p) in train mode; `state` = glue.DropoutState (seed + device step counter) or a false value = off."""
    return G.dropout2d(x, c or x.shape[3], p, state) if (state and p > 0) else x


def _sft_like(P, p, feats, kvec):
    """SFTLikeBlock.forward (model/modeling/blocks.py:105-120) on cat(features[64], kernel map[441]); the spatially constant
    kernel-map half of conv 0 is evaluated on a 3x3 image of border classes (as in _sft)."""
    n, h, w, fc = feats.shape
    cond = _expand_vec(kvec, 3, 3)

    def branch(name):
        bp = p + ".conv_%s" % name
        w0 = P[bp + ".0.layer.weight"]
        fcr = w0.shape[1] - kvec.shape[1]
        t = conv2d(feats, w0, None, padding=1, cin_range=(0, fcr))
        tb = conv2d(cond, w0, P[bp + ".0.layer.bias"], padding=1, cin_range=(fcr, kvec.shape[1]))
        t = prelu(G.add(t, _expand_classes(tb, h, w, 1)), P[bp + ".0.act.weight"])
        return conv2d(t, P[bp + ".1.layer.weight"], P[bp + ".1.layer.bias"], padding=1)
    return G.sft_combine(feats, branch("scale"), branch("shift"))


def pspnet_forward(P, img, sizes=(1, 2, 3, 6), prefix="segmentation_model.", bn_training=True, dropout=True, kvec=None):
    """img fp32 [B,3,H,W] (already normalised) -> (seg, aux) fp32 [B,1,H,W].  With `kvec` (B, 441) and blur_skip
    parameters: PSPNet_BlurSkip.forward (pspnet.py:174-207), p + BlurSkip(p, kernel)."""
    p = prefix
    H, W = img.shape[2:]
    x = to_nhwc(img)
    x = _bn(P, p + "feats.bn1", conv2d(x, P[p + "feats.conv1.weight"], None, stride=2, padding=3), bn_training, relu=True)
    x = G.maxpool3s2(x)
    x3 = None
    for li, (planes, blocks, stride, dil) in enumerate(RESNET34_LAYERS, 1):
        for b in range(blocks):
            bp = p + "feats.layer%d.%d" % (li, b)
            st = stride if b == 0 else 1
            d = 1 if b == 0 else dil
            out = _bn(P, bp + ".bn1", conv2d(x, P[bp + ".conv1.weight"], None, stride=st, padding=d, dilation=d), bn_training,
                      relu=True)
            res = x
            if (bp + ".downsample.0.weight") in P:
                res = _bn(P, bp + ".downsample.1", conv2d(x, P[bp + ".downsample.0.weight"], None, stride=st), bn_training)
            x = _bn(P, bp + ".bn2", conv2d(out, P[bp + ".conv2.weight"], None, padding=d, dilation=d), bn_training,
                    relu=True, res=res)                                # relu(bn2(conv2) + residual), extractors.py:62-70
        if li == 3:
            x3 = x
    f = x
    h, w = f.shape[1:3]
    priors = []
    for i, s in enumerate(sizes):
        pooled = G.adaptive_avgpool(f, s)
        pr = conv2d(pooled, P[p + "psp.stages.%d.1.weight" % i], None)
        priors.append(G.bilinear(pr, (h, w)))
    priors.append(f)
    y = conv2d(_cat(priors), P[p + "psp.bottleneck.weight"], P[p + "psp.bottleneck.bias"], act="relu")
    y = _drop(y, 0.3, dropout)
    for name, dp in (("up_1", 0.15), ("up_2", 0.15), ("up_3", 0.15)):      # drop_2 after every up block (pspnet.py:106-113)
        y = G.bilinear(y, (2 * y.shape[1], 2 * y.shape[2]))
        y = _bn(P, p + name + ".conv.1", conv2d(y, P[p + name + ".conv.0.weight"], P[p + name + ".conv.0.bias"], padding=1),
                bn_training)
        y = prelu(y, P[p + name + ".conv.2.weight"])
        y = _drop(y, dp, dropout)
    if kvec is not None:
        t, i = y, 0
        while (p + "blur_skip.%d.conv_scale.0.layer.weight" % (2 * i)) in P:
            t = _sft_like(P, p + "blur_skip.%d" % (2 * i), t, kvec)
            bp = p + "blur_skip.%d" % (2 * i + 1)
            t = _bn(P, bp + ".norm", conv2d(t, P[bp + ".layer.weight"], None, padding=1), bn_training, relu=True)
            i += 1
        y = G.add(y, t)
    seg = torch.sigmoid(to_nchw(conv2d(y, P[p + "final.0.weight"], P[p + "final.0.bias"]), 1))
    a = _bn(P, p + "aux.1", conv2d(x3, P[p + "aux.0.weight"], None, padding=1), bn_training, relu=True)
    a = _drop(a, 0.1, dropout)
    a = torch.sigmoid(to_nchw(conv2d(a, P[p + "aux.4.weight"], P[p + "aux.4.bias"]), 1))
    aux = F.interpolate(a, size=(H, W), mode="bilinear", align_corners=True)
    _flush_tracked()
    return seg, aux


# ---------------------------------------------------------------------------------------------- HRNet-W48 + OCR (train mode)
def _cbr(P, pc, pb, x, training, stride=1, padding=0, relu=True, res=None):
    return _bn(P, pb, conv2d(x, P[pc + ".weight"], P.get(pc + ".bias"), stride=stride, padding=padding), training,
               relu=relu, res=res)


def _resize_ac(x, size):
    return G.bilinear(x, size, align_corners=True)


def hrnet_w48(P, p, x, training):
    """HighResolutionNet.forward (hrnet_ocr/backbones/hrnet/hrnet_backbone.py:514-572) on NHWC tensors."""
    from .params import HRNET48_STAGES
    x = _cbr(P, p + "conv1", p + "bn1", x, training, stride=2, padding=1)
    x = _cbr(P, p + "conv2", p + "bn2", x, training, stride=2, padding=1)
    for i in range(4):
        bp = p + "layer1.%d" % i
        out = _cbr(P, bp + ".conv1", bp + ".bn1", x, training)
        out = _cbr(P, bp + ".conv2", bp + ".bn2", out, training, padding=1)
        res = _cbr(P, bp + ".downsample.0", bp + ".downsample.1", x, training, relu=False) if i == 0 else x
        x = _cbr(P, bp + ".conv3", bp + ".bn3", out, training, relu=True, res=res)
    ys, pre = [x], (256,)
    for si, (modules, chans) in enumerate(HRNET48_STAGES, 2):
        t = p + "transition%d" % (si - 1)
        xs = []
        for i, c in enumerate(chans):
            if i < len(pre):
                xs.append(_cbr(P, t + ".%d.0" % i, t + ".%d.1" % i, ys[i], training, padding=1) if c != pre[i] else ys[i])
            else:
                xs.append(_cbr(P, t + ".%d.0.0" % i, t + ".%d.0.1" % i, ys[-1], training, stride=2, padding=1))
        for m in range(modules):
            mp = p + "stage%d.%d" % (si, m)
            for bi in range(len(chans)):
                for k in range(4):
                    bp = mp + ".branches.%d.%d" % (bi, k)
                    out = _cbr(P, bp + ".conv1", bp + ".bn1", xs[bi], training, padding=1)
                    xs[bi] = _cbr(P, bp + ".conv2", bp + ".bn2", out, training, padding=1, relu=True, res=xs[bi])
            fused = []
            for i in range(len(chans)):
                y = None
                for j in range(len(chans)):
                    fp = mp + ".fuse_layers.%d.%d" % (i, j)
                    if j == i:
                        term = xs[j]
                    elif j > i:
                        term = _resize_ac(_cbr(P, fp + ".0", fp + ".1", xs[j], training, relu=False), xs[i].shape[1:3])
                    else:
                        term = xs[j]
                        for k in range(i - j):
                            term = _cbr(P, fp + ".%d.0" % k, fp + ".%d.1" % k, term, training, stride=2, padding=1,
                                        relu=(k != i - j - 1))
                    y = term if y is None else G.add(y, term)
                fused.append(G.relu(y))
            xs = fused
        ys, pre = xs, chans
    return ys, HRNET48_STAGES[-1][1]


def hrnet_ocr_forward(P, img, prefix="segmentation_model.", bn_training=True, dropout=True):
    """HRNet_W48_OCR.forward (hrnet_ocr/nets/hrnet.py:137-158) in train mode; SpatialGather_Module / _ObjectAttentionBlock /
    SpatialOCR_Module (modules/spatial_ocr_block.py:49-66, 172-214, 281-303).  img fp32 [B,3,H,W] -> (seg, aux) fp32."""
    p = prefix
    tr = bn_training
    H, W = img.shape[2:]
    ys, chans = hrnet_w48(P, p + "backbone.", to_nhwc(img), tr)
    h, w = ys[0].shape[1:3]
    feats = _cat([ys[0]] + [_resize_ac(y, (h, w)) for y in ys[1:]], real=chans)                 # 720 -> 768
    a = _bn(P, p + "aux_head.1.0", conv2d(feats, P[p + "aux_head.0.weight"], P[p + "aux_head.0.bias"], padding=1), tr, relu=True)
    out_aux = to_nchw(conv2d(a, P[p + "aux_head.2.weight"], P[p + "aux_head.2.bias"]), 1)      # [B,1,h,w] fp32
    f = _bn(P, p + "conv3x3.1.0", conv2d(feats, P[p + "conv3x3.0.weight"], P[p + "conv3x3.0.bias"], padding=1), tr, relu=True)
    B, C = f.shape[0], f.shape[3]
    probs = F.softmax(out_aux.view(B, 1, -1), dim=2)                                            # (B, K=1, HW)
    ctx = torch.matmul(probs, f.view(B, h * w, C).float())                                      # (B, 1, C)
    ctx = ctx.to(torch.bfloat16).view(B, 1, 1, C)                                               # object region feature
    o = p + "ocr_distri_head.object_context_block."

    def seq2(name, t):
        t = _bn(P, o + name + ".1.0", conv2d(t, P[o + name + ".0.weight"], P.get(o + name + ".0.bias")), tr, relu=True)
        return _bn(P, o + name + ".3.0", conv2d(t, P[o + name + ".2.weight"], P.get(o + name + ".2.bias")), tr, relu=True)

    query = seq2("f_pixel", f).view(B, h * w, 256).float()
    key = seq2("f_object", ctx).view(B, 1, 256).float().permute(0, 2, 1)
    value = _bn(P, o + "f_down.1.0", conv2d(ctx, P[o + "f_down.0.weight"], P.get(o + "f_down.0.bias")), tr, relu=True)
    sim = F.softmax((256 ** -0.5) * torch.matmul(query, key), dim=-1)                           # (B, HW, K=1) == 1
    context = torch.matmul(sim, value.view(B, 1, 256).float()).to(torch.bfloat16).view(B, h, w, 256)
    context = _bn(P, o + "f_up.1.0", conv2d(context, P[o + "f_up.0.weight"], P.get(o + "f_up.0.bias")), tr, relu=True)
    q = p + "ocr_distri_head.conv_bn_dropout."
    f2 = _bn(P, q + "1.0", conv2d(_cat((context, f)), P[q + "0.weight"], P.get(q + "0.bias")), tr, relu=True)
    f2 = _drop(f2, 0.05, dropout)
    out = to_nchw(conv2d(f2, P[p + "cls_head.weight"], P[p + "cls_head.bias"]), 1)
    up = lambda t: F.interpolate(t, size=(H, W), mode="bilinear", align_corners=True)
    _flush_tracked()
    return torch.sigmoid(up(out)), torch.sigmoid(up(out_aux))
