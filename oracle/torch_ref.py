"""TEST INFRASTRUCTURE ONLY -- plain PyTorch fp32 restatement of the reference's eval forward
(JointModel: KBPN blind SR -> clip -> instance norm -> PSPNet), written functionally over a state_dict.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
It is pinned against the UNMODIFIED reference by tests/golden/gen_golden.py (run where /root/reference
exists): the committed fixtures hold outputs of the real `JointModel` on the same synthetic weights.

Each function cites the reference lines it restates (paths relative to the reference root).
"""
import torch
import torch.nn.functional as F


def _conv(sd, p, x, stride=1, padding=0, dilation=1):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding, dilation=dilation)


def _convblock(sd, p, x, k, stride=1, padding=0, act=None):
    """ConvBlock: conv -> (no norm) -> activation; model/modeling/kbpn.py:266-270, 190-248."""
    y = F.conv2d(x, sd[p + ".layer.weight"], sd.get(p + ".layer.bias"), stride=stride, padding=padding)
    return _act(sd, p, y, act)


def _deconvblock(sd, p, x, act="prelu"):
    """DeconvBlock 8/4/2: model/modeling/kbpn.py:273-277."""
    y = F.conv_transpose2d(x, sd[p + ".layer.weight"], sd.get(p + ".layer.bias"), stride=4, padding=2)
    return _act(sd, p, y, act)


def _act(sd, p, y, act):
    if act == "prelu":
        return F.prelu(y, sd[p + ".act.weight"])
    if act == "relu":
        return F.relu(y)
    if act == "lrelu":
        return F.leaky_relu(y, 0.01)
    assert act is None
    return y


def _gap(x):
    return x.mean(dim=(2, 3), keepdim=True)


def _upscale_kernel(vec, k_out):
    """7x7 -> 21x21 bicubic (non-antialiased, align_corners=False): kbpn.py:335-341, 580-602."""
    k = int(round(vec.shape[1] ** 0.5))
    ker = F.interpolate(vec.view(vec.shape[0], 1, k, k), size=(k_out, k_out), mode="bicubic")
    return ker


def predictor_with_gap(sd, p, x, k_out):
    """predictor_withGAP.forward: kbpn.py:320-341 -> (B, k_out^2, 1, 1), normalised after upsampling."""
    z = x
    for i in range(3):
        z = _convblock(sd, p + ".feat_ext.%d" % i, z, 3, padding=1, act="prelu")
    vec = _gap(z)
    ker = _upscale_kernel(vec, k_out)
    ker = ker / ker.sum(dim=(2, 3), keepdim=True)
    return ker.view(ker.shape[0], k_out * k_out, 1, 1)


def kernel_predictor_ikc(sd, p, sr, pre_kernel_vec, k_out):
    """KernelPredictorLikeIKC.forward: kbpn.py:562-578. `pre_kernel_vec` is (B,441,1,1) (spatially constant)."""
    H, W = sr.shape[2:]
    fsr = _convblock(sd, p + ".fe_SR.0", sr, 3, padding=1, act="relu")
    fsr = _convblock(sd, p + ".fe_SR.1", fsr, 1, act="lrelu")
    fsr = _convblock(sd, p + ".fe_SR.2", fsr, 3, padding=1, act="lrelu")
    fsr = _convblock(sd, p + ".fe_SR.3", fsr, 3, padding=1, act="lrelu")
    fsr = _convblock(sd, p + ".fe_SR.4", fsr, 3, padding=1, act="lrelu")
    fh = pre_kernel_vec.expand(-1, -1, H, W)
    fh = _convblock(sd, p + ".fe_kernel.0", fh, 3, padding=1, act="lrelu")
    fh = _convblock(sd, p + ".fe_kernel.1", fh, 3, padding=1, act="lrelu")
    d = torch.cat((fsr, fh), dim=1)
    d = _convblock(sd, p + ".fe_cat.0", d, 1, act="lrelu")
    d = _convblock(sd, p + ".fe_cat.1", d, 3, padding=1, act="lrelu")
    d = _convblock(sd, p + ".fe_cat.2", d, 3, padding=1, act=None)
    delta = _upscale_kernel(_gap(d), k_out).view(sr.shape[0], k_out * k_out, 1, 1)
    return pre_kernel_vec + delta


def up_block(sd, p, x):
    """UpBlock.forward: kbpn.py:464-469."""
    x = _convblock(sd, p + ".conv", x, 1, act="prelu")
    h0 = _deconvblock(sd, p + ".up_conv1", x)
    l0 = _convblock(sd, p + ".up_conv2", h0, 8, stride=4, padding=2, act="prelu")
    h1 = _deconvblock(sd, p + ".up_conv3", l0 - x)
    return h1 + h0


def down_block(sd, p, x):
    """DownBlock.forward: kbpn.py:484-489."""
    x = _convblock(sd, p + ".conv", x, 1, act="prelu")
    l0 = _convblock(sd, p + ".down_conv1", x, 8, stride=4, padding=2, act="prelu")
    h0 = _deconvblock(sd, p + ".down_conv2", l0)
    l1 = _convblock(sd, p + ".down_conv3", h0 - x, 8, stride=4, padding=2, act="prelu")
    return l1 + l0


def k_block(sd, p, concat_h, h, x_lr, kvec, k_out, scale, predict_kernel=True):
    """KBlock.forward (SUM_LR_ERROR_POS='HR'): kbpn.py:382-412; the kernel predictor is skipped during SR pretrain (:386)."""
    sr_t = _convblock(sd, p + ".sr_reconst", concat_h, 3, padding=1, act=None)
    d_kernel = kernel_predictor_ikc(sd, p + ".kernel_predictor", sr_t, kvec, k_out) if predict_kernel else kvec
    vec = d_kernel / d_kernel.sum(dim=1).view(-1, 1, 1, 1)       # GAP of a constant map is the vector itself
    weight = vec.view(-1, 1, k_out, k_out)
    pad = (k_out - 1) // 2
    outs = []
    for k in range(sr_t.shape[0]):                                # per-sample depthwise blur, stride 4 (:395-402)
        outs.append(F.conv2d(sr_t[k:k + 1], weight[k].expand(3, 1, k_out, k_out), stride=scale, padding=pad, groups=3))
    pseudo_lr = torch.cat(outs, dim=0)
    e_h = _deconvblock(sd, p + ".up_conv1", pseudo_lr - x_lr)
    return h + e_h, vec, sr_t


def sft_layer(sd, p, features, kvec):
    """SFTlayer.forward: kbpn.py:511-518."""
    cond = kvec.expand(-1, -1, features.shape[2], features.shape[3])
    c = torch.cat((features, cond), dim=1)
    scale = _conv(sd, p + ".SFT_scale_conv1", F.leaky_relu(_conv(sd, p + ".SFT_scale_conv0", c, padding=1), 0.1), padding=1)
    shift = _conv(sd, p + ".SFT_shift_conv1", F.leaky_relu(_conv(sd, p + ".SFT_shift_conv0", c, padding=1), 0.1), padding=1)
    return features * torch.sigmoid(scale) + shift


def kbpn_forward(sd, x, num_stages=4, k_out=21, scale=4, prefix="sr_model.", return_intermediates=False, gt_kernel=None):
    """KBPN.forward with iter=-1 (eval): kbpn.py:84-116. Returns sr (B,3,4h,4w) and kernel vec (B,441,1,1)."""
    p = prefix
    f = x
    for i in (0, 2, 4, 6):                                        # VGG16 head, :42-44
        f = F.relu(_conv(sd, p + "feat.%d" % i, f, padding=1))
    init_f = f
    if gt_kernel is not None:                                     # SR-module pre-training: kbpn.py:89-91
        kvec = gt_kernel.reshape(-1, k_out * k_out, 1, 1)
    else:
        kvec = predictor_with_gap(sd, p + "predictor", init_f, k_out)
    inter = {"init_f": init_f, "init_kernel": kvec}
    low, concat_h, concat_l = init_f, None, None
    for s in range(num_stages):                                   # KernelBackProjectionStageWithSFT.forward :172-189
        sp = p + "back_projection_stages.%d" % s
        h = up_block(sd, sp + ".up", low)
        pre = h if concat_h is None else torch.cat((concat_h, h), dim=1)
        h, kvec, sr_t = k_block(sd, sp + ".kb", pre, h, x, kvec, k_out, scale, predict_kernel=gt_kernel is None)
        inter["sr_t%d" % s] = sr_t
        inter["kvec%d" % s] = kvec
        concat_h = h if concat_h is None else torch.cat((concat_h, h), dim=1)
        if s < num_stages - 1:
            low = down_block(sd, sp + ".down", concat_h)
            concat_l = low if concat_l is None else torch.cat((concat_l, low), dim=1)
            low = sft_layer(sd, sp + ".sft", concat_l, kvec)
    sr = _convblock(sd, p + "output_conv", concat_h, 3, padding=1, act=None)
    sr = sr + F.interpolate(x, scale_factor=scale, mode="bicubic")   # nn.Upsample(scale_factor=4,'bicubic') :70,113
    if return_intermediates:
        return sr, kvec, inter
    return sr, kvec


BN_TRAIN = False      # oracle/train_ref.py switches this on: BatchNorm uses batch statistics (model.train())


def _bn(sd, p, x):
    if BN_TRAIN:
        return F.batch_norm(x, None, None, sd[p + ".weight"], sd[p + ".bias"], training=True, eps=1e-5)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training=False, eps=1e-5)


RESNET34_LAYERS = ((64, 3, 1, 1), (128, 4, 2, 1), (256, 6, 1, 2), (512, 3, 1, 4))


def resnet34_dilated(sd, p, x):
    """ResNet.forward + BasicBlock.forward: pspnet_pytorch/extractors.py:150-161, 52-70; the first block of a
    layer is built with the layer stride and dilation 1, later blocks with the layer dilation (:143-146)."""
    x = F.relu(_bn(sd, p + "bn1", _conv(sd, p + "conv1", x, stride=2, padding=3)))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    x3 = None
    for li, (planes, blocks, stride, dil) in enumerate(RESNET34_LAYERS, 1):
        for b in range(blocks):
            bp = p + "layer%d.%d" % (li, b)
            st = stride if b == 0 else 1
            d = 1 if b == 0 else dil
            out = F.relu(_bn(sd, bp + ".bn1", _conv(sd, bp + ".conv1", x, stride=st, padding=d, dilation=d)))
            out = _bn(sd, bp + ".bn2", _conv(sd, bp + ".conv2", out, padding=d, dilation=d))
            res = x
            if (bp + ".downsample.0.weight") in sd:
                res = _bn(sd, bp + ".downsample.1", _conv(sd, bp + ".downsample.0", x, stride=st))
            x = F.relu(out + res)
        if li == 3:
            x3 = x
    return x, x3


def _sft_like_block(sd, p, feats, cond):
    """SFTLikeBlock.forward: model/modeling/blocks.py:105-120 (PReLU after conv 0, sigmoid / none after conv 1)."""
    c = torch.cat((feats, cond), 1)
    def branch(name, last):
        t = F.prelu(_conv(sd, p + ".conv_%s.0.layer" % name, c, padding=1), sd[p + ".conv_%s.0.act.weight" % name])
        t = _conv(sd, p + ".conv_%s.1.layer" % name, t, padding=1)
        return torch.sigmoid(t) if last == "sigmoid" else t
    return feats * branch("scale", "sigmoid") + branch("shift", None)


def pspnet_forward(sd, x, prefix="segmentation_model.", sizes=(1, 2, 3, 6), kvec=None):
    """PSPNet.forward in eval mode (dropout = identity): pspnet_pytorch/pspnet.py:95-123, 23-57.
    With `kvec` (B,441,1,1) and blur_skip weights: PSPNet_BlurSkip.forward, pspnet.py:174-207."""
    p = prefix
    H, W = x.shape[2:]
    f, x3 = resnet34_dilated(sd, p + "feats.", x)
    h, w = f.shape[2:]
    priors = []
    for i, s in enumerate(sizes):
        pr = F.conv2d(F.adaptive_avg_pool2d(f, (s, s)), sd[p + "psp.stages.%d.1.weight" % i])
        priors.append(F.interpolate(pr, size=(h, w), mode="bilinear"))
    priors.append(f)
    y = F.relu(_conv(sd, p + "psp.bottleneck", torch.cat(priors, 1)))
    for name in ("up_1", "up_2", "up_3"):
        y = F.interpolate(y, size=(2 * y.shape[2], 2 * y.shape[3]), mode="bilinear")
        y = _bn(sd, p + name + ".conv.1", _conv(sd, p + name + ".conv.0", y, padding=1))
        y = F.prelu(y, sd[p + name + ".conv.2.weight"])
    if kvec is not None:
        cond = kvec.expand(-1, -1, H, W)
        t = y
        i = 0
        while (p + "blur_skip.%d.conv_scale.0.layer.weight" % (2 * i)) in sd:
            t = _sft_like_block(sd, p + "blur_skip.%d" % (2 * i), t, cond)
            bp = p + "blur_skip.%d" % (2 * i + 1)
            t = F.relu(F.batch_norm(_conv(sd, bp + ".layer", t, padding=1), sd[bp + ".norm.running_mean"],
                                    sd[bp + ".norm.running_var"], sd[bp + ".norm.weight"], sd[bp + ".norm.bias"],
                                    training=False, eps=1e-5))
            i += 1
        y = y + t
    seg = torch.sigmoid(_conv(sd, p + "final.0", y))
    a = F.relu(_bn(sd, p + "aux.1", _conv(sd, p + "aux.0", x3, padding=1)))
    a = torch.sigmoid(_conv(sd, p + "aux.4", a))
    aux = F.interpolate(a, size=(H, W), mode="bilinear", align_corners=True)
    return seg, aux


HRNET48_STAGES = ((1, (48, 96)), (4, (48, 96, 192)), (3, (48, 96, 192, 384)))


def _cbr(sd, p_conv, p_bn, x, stride=1, padding=0, relu=True):
    y = _bn(sd, p_bn, _conv(sd, p_conv, x, stride=stride, padding=padding))
    return F.relu(y) if relu else y


def hrnet_w48(sd, p, x):
    """HighResolutionNet.forward: hrnet_ocr/backbones/hrnet/hrnet_backbone.py:514-572 (stem :308-318, Bottleneck layer1,
    transitions :402-447, HighResolutionModule.forward :265-290 with bilinear align_corners=True fusion)."""
    x = _cbr(sd, p + "conv1", p + "bn1", x, stride=2, padding=1)
    x = _cbr(sd, p + "conv2", p + "bn2", x, stride=2, padding=1)
    for i in range(4):
        bp = p + "layer1.%d" % i
        out = _cbr(sd, bp + ".conv1", bp + ".bn1", x)
        out = _cbr(sd, bp + ".conv2", bp + ".bn2", out, padding=1)
        out = _cbr(sd, bp + ".conv3", bp + ".bn3", out, relu=False)
        res = _cbr(sd, bp + ".downsample.0", bp + ".downsample.1", x, relu=False) if i == 0 else x
        x = F.relu(out + res)
    ys = [x]
    pre = (256,)
    for si, (modules, chans) in enumerate(HRNET48_STAGES, 2):
        t = p + "transition%d" % (si - 1)
        xs = []
        for i, c in enumerate(chans):
            if i < len(pre):
                xs.append(_cbr(sd, t + ".%d.0" % i, t + ".%d.1" % i, ys[i], padding=1) if c != pre[i] else ys[i])
            else:
                xs.append(_cbr(sd, t + ".%d.0.0" % i, t + ".%d.0.1" % i, ys[-1], stride=2, padding=1))
        for m in range(modules):
            mp = p + "stage%d.%d" % (si, m)
            for bi in range(len(chans)):
                for k in range(4):
                    bp = mp + ".branches.%d.%d" % (bi, k)
                    out = _cbr(sd, bp + ".conv1", bp + ".bn1", xs[bi], padding=1)
                    out = _cbr(sd, bp + ".conv2", bp + ".bn2", out, padding=1, relu=False)
                    xs[bi] = F.relu(out + xs[bi])
            fused = []
            for i in range(len(chans)):
                y = None
                for j in range(len(chans)):
                    fp = mp + ".fuse_layers.%d.%d" % (i, j)
                    if j == i:
                        term = xs[j]
                    elif j > i:
                        term = F.interpolate(_cbr(sd, fp + ".0", fp + ".1", xs[j], relu=False), size=xs[i].shape[-2:],
                                             mode="bilinear", align_corners=True)
                    else:
                        term = xs[j]
                        for k in range(i - j):
                            term = _cbr(sd, fp + ".%d.0" % k, fp + ".%d.1" % k, term, stride=2, padding=1, relu=(k != i - j - 1))
                    y = term if y is None else y + term
                fused.append(F.relu(y))
            xs = fused
        ys = xs
        pre = chans
    return ys


def hrnet_ocr_forward(sd, x, prefix="segmentation_model."):
    """HRNet_W48_OCR.forward in eval mode: hrnet_ocr/nets/hrnet.py:137-158; SpatialGather_Module
    modules/spatial_ocr_block.py:49-66; _ObjectAttentionBlock.forward :172-196; SpatialOCR_Module.forward :281-303."""
    p = prefix
    H, W = x.shape[2:]
    ys = hrnet_w48(sd, p + "backbone.", x)
    h, w = ys[0].shape[2:]
    feats = torch.cat([ys[0]] + [F.interpolate(y, size=(h, w), mode="bilinear", align_corners=True) for y in ys[1:]], 1)
    a = F.relu(_bn(sd, p + "aux_head.1.0", _conv(sd, p + "aux_head.0", feats, padding=1)))
    out_aux = _conv(sd, p + "aux_head.2", a)
    f = F.relu(_bn(sd, p + "conv3x3.1.0", _conv(sd, p + "conv3x3.0", feats, padding=1)))
    B, C = f.shape[:2]
    probs = F.softmax(out_aux.view(B, out_aux.shape[1], -1), dim=2)                       # (B, K=1, HW)
    ctx = torch.matmul(probs, f.view(B, C, -1).permute(0, 2, 1)).permute(0, 2, 1).unsqueeze(3)   # (B, C, K, 1)
    o = p + "ocr_distri_head.object_context_block."
    def seq2(name, t):
        t = F.relu(_bn(sd, o + name + ".1.0", _conv(sd, o + name + ".0", t)))
        return F.relu(_bn(sd, o + name + ".3.0", _conv(sd, o + name + ".2", t)))
    query = seq2("f_pixel", f).view(B, 256, -1).permute(0, 2, 1)
    key = seq2("f_object", ctx).view(B, 256, -1)
    value = F.relu(_bn(sd, o + "f_down.1.0", _conv(sd, o + "f_down.0", ctx))).view(B, 256, -1).permute(0, 2, 1)
    sim = F.softmax((256 ** -0.5) * torch.matmul(query, key), dim=-1)
    context = torch.matmul(sim, value).permute(0, 2, 1).contiguous().view(B, 256, h, w)
    context = F.relu(_bn(sd, o + "f_up.1.0", _conv(sd, o + "f_up.0", context)))
    q = p + "ocr_distri_head.conv_bn_dropout."
    f2 = F.relu(_bn(sd, q + "1.0", _conv(sd, q + "0", torch.cat([context, f], 1))))
    out = _conv(sd, p + "cls_head", f2)
    up = lambda t: F.interpolate(t, size=(H, W), mode="bilinear", align_corners=True)
    return torch.sigmoid(up(out)), torch.sigmoid(up(out_aux))


def joint_forward(sd, x, k_out=21, num_stages=4, blur_skip=False, hrnet=False):
    """JointModel.forward (KBPN + PSPNet, NORM_SR_OUTPUT='instance'): model/modeling/build_model.py:466-496,
    clip_sr :143-146, norm_sr :135-137 (fresh InstanceNorm2d(3): eps 1e-5, biased variance, no affine)."""
    sr, kvec = kbpn_forward(sd, x, num_stages=num_stages, k_out=k_out)
    sr = sr.clamp(0.0, 1.0)
    if hrnet:
        seg, aux = hrnet_ocr_forward(sd, F.instance_norm(sr, eps=1e-5))
    else:
        seg, aux = pspnet_forward(sd, F.instance_norm(sr, eps=1e-5), kvec=kvec if blur_skip else None)
    kp = kvec / kvec.sum(dim=1).view(-1, 1, 1, 1)
    return sr, seg, kp.view(-1, 1, k_out, k_out), aux
