"""Parameter inventory of the CSBSR networks: names and shapes identical to the reference's state_dict
(SURVEY.md App. A.4), so released / reference checkpoints load with strict=True.

KBPN:   model/modeling/kbpn.py:17-82 (KBPN.__init__), :157-170 (stage), :344-375 (KBlock), :450-489
        (Up/DownBlock), :493-507 (SFTlayer), :521-541 (KernelPredictorLikeIKC), :292-312 (predictor_withGAP)
PSPNet: model/modeling/pspnet_pytorch/pspnet.py:23-86, extractors.py:37-148 (ResNet-34, BasicBlock)
"""
import zlib
from collections import OrderedDict

import torch
from torch import nn


def kbpn_param_shapes(num_stages=4, md_ch=128, k_est=7, k_out=21, num_ch=3, reduction_ch=32):
    """OrderedDict name -> shape for `sr_model.*` (KBPN, scale 4: conv setting 8/4/2)."""
    P = OrderedDict()
    kc = k_est * k_est          # 49
    cond = k_out * k_out        # 441

    def convblock(pfx, cin, cout, k, bias=False, act=None):
        P[pfx + ".layer.weight"] = (cout, cin, k, k)
        if bias:
            P[pfx + ".layer.bias"] = (cout,)
        if act == "prelu":
            P[pfx + ".act.weight"] = (1,)

    def deconvblock(pfx, cin, cout, k, act="prelu"):
        P[pfx + ".layer.weight"] = (cin, cout, k, k)
        if act == "prelu":
            P[pfx + ".act.weight"] = (1,)

    for idx, (ci, co) in zip((0, 2, 4, 6), ((num_ch, 64), (64, 64), (64, 128), (128, 128))):
        P["feat.%d.weight" % idx] = (co, ci, 3, 3)
        P["feat.%d.bias" % idx] = (co,)
    for i in range(3):
        convblock("predictor.feat_ext.%d" % i, md_ch, md_ch if i < 2 else kc, 3, act="prelu")
    for s in range(num_stages):
        stages = s + 1
        up_stages = stages - 1 if stages > 1 else 1
        sp = "back_projection_stages.%d" % s
        convblock(sp + ".up.conv", md_ch * up_stages, md_ch, 1, bias=True, act="prelu")
        convblock(sp + ".up.up_conv2", md_ch, md_ch, 8, act="prelu")
        deconvblock(sp + ".up.up_conv1", md_ch, md_ch, 8)
        deconvblock(sp + ".up.up_conv3", md_ch, md_ch, 8)
        convblock(sp + ".kb.sr_reconst", stages * md_ch, 3, 3)
        kp = sp + ".kb.kernel_predictor"
        convblock(kp + ".fe_SR.0", 3, kc, 3)
        convblock(kp + ".fe_SR.1", kc, reduction_ch, 1)
        convblock(kp + ".fe_SR.2", reduction_ch, reduction_ch, 3)
        convblock(kp + ".fe_SR.3", reduction_ch, reduction_ch, 3)
        convblock(kp + ".fe_SR.4", reduction_ch, kc, 3)
        convblock(kp + ".fe_kernel.0", cond, kc, 3)
        convblock(kp + ".fe_kernel.1", kc, kc, 3)
        convblock(kp + ".fe_cat.0", 2 * kc, reduction_ch, 1)
        convblock(kp + ".fe_cat.1", reduction_ch, reduction_ch, 3)
        convblock(kp + ".fe_cat.2", reduction_ch, kc, 3)
        deconvblock(sp + ".kb.up_conv1", 3, md_ch, 8)
        if s < num_stages - 1:
            convblock(sp + ".down.conv", md_ch * stages, md_ch, 1, bias=True, act="prelu")
            convblock(sp + ".down.down_conv1", md_ch, md_ch, 8, act="prelu")
            convblock(sp + ".down.down_conv3", md_ch, md_ch, 8, act="prelu")
            deconvblock(sp + ".down.down_conv2", md_ch, md_ch, 8)
            cc = stages * md_ch + cond
            for br in ("scale", "shift"):
                P[sp + ".sft.SFT_%s_conv0.weight" % br] = (cc, cc, 3, 3)
                P[sp + ".sft.SFT_%s_conv0.bias" % br] = (cc,)
                P[sp + ".sft.SFT_%s_conv1.weight" % br] = (stages * md_ch, cc, 3, 3)
                P[sp + ".sft.SFT_%s_conv1.bias" % br] = (stages * md_ch,)
    convblock("output_conv", num_stages * md_ch, num_ch, 3)
    return P


def _bn(P, pfx, c):
    P[pfx + ".weight"] = (c,)
    P[pfx + ".bias"] = (c,)
    P[pfx + ".running_mean"] = (c,)
    P[pfx + ".running_var"] = (c,)
    P[pfx + ".num_batches_tracked"] = ()


RESNET34_LAYERS = ((64, 3, 1, 1), (128, 4, 2, 1), (256, 6, 1, 2), (512, 3, 1, 4))   # planes, blocks, stride, dilation


def pspnet_param_shapes(n_classes=1, sizes=(1, 2, 3, 6), psp_size=512, deep_features_size=256, blur_dim=None,
                        n_layer_blurskip=2):
    """OrderedDict name -> shape for `segmentation_model.*` (PSPNet on a dilated ResNet-34); with `blur_dim`
    (= KERNEL_SIZE_OUTPUT**2) also the BlurSkip blocks of PSPNet_BlurSkip (pspnet.py:127-160, blocks.py:105-120)."""
    P = OrderedDict()
    P["feats.conv1.weight"] = (64, 3, 7, 7)
    _bn(P, "feats.bn1", 64)
    inplanes = 64
    for li, (planes, blocks, stride, _dil) in enumerate(RESNET34_LAYERS, 1):
        for b in range(blocks):
            bp = "feats.layer%d.%d" % (li, b)
            P[bp + ".conv1.weight"] = (planes, inplanes if b == 0 else planes, 3, 3)
            _bn(P, bp + ".bn1", planes)
            P[bp + ".conv2.weight"] = (planes, planes, 3, 3)
            _bn(P, bp + ".bn2", planes)
            if b == 0 and (stride != 1 or inplanes != planes):
                P[bp + ".downsample.0.weight"] = (planes, inplanes, 1, 1)
                _bn(P, bp + ".downsample.1", planes)
        inplanes = planes
    for i, _ in enumerate(sizes):
        P["psp.stages.%d.1.weight" % i] = (psp_size, psp_size, 1, 1)
    P["psp.bottleneck.weight"] = (1024, psp_size * (len(sizes) + 1), 1, 1)
    P["psp.bottleneck.bias"] = (1024,)
    for name, (ci, co) in (("up_1", (1024, 256)), ("up_2", (256, 64)), ("up_3", (64, 64))):
        P[name + ".conv.0.weight"] = (co, ci, 3, 3)
        P[name + ".conv.0.bias"] = (co,)
        _bn(P, name + ".conv.1", co)
        P[name + ".conv.2.weight"] = (1,)
    if blur_dim is not None:
        cc = blur_dim + 64
        for i in range(n_layer_blurskip):
            for br in ("scale", "shift"):
                bp = "blur_skip.%d.conv_%s" % (2 * i, br)
                P[bp + ".0.layer.weight"] = (cc, cc, 3, 3)
                P[bp + ".0.layer.bias"] = (cc,)
                P[bp + ".0.act.weight"] = (1,)
                P[bp + ".1.layer.weight"] = (64, cc, 3, 3)
                P[bp + ".1.layer.bias"] = (64,)
            P["blur_skip.%d.layer.weight" % (2 * i + 1)] = (64, 64, 3, 3)
            _bn(P, "blur_skip.%d.norm" % (2 * i + 1), 64)
    P["final.0.weight"] = (n_classes, 64, 1, 1)
    P["final.0.bias"] = (n_classes,)
    P["aux.0.weight"] = (256, deep_features_size, 3, 3)
    _bn(P, "aux.1", 256)
    P["aux.4.weight"] = (n_classes, 256, 1, 1)
    P["aux.4.bias"] = (n_classes,)
    return P


HRNET48_STAGES = (          # (modules, channels per branch): hrnet_ocr/backbones/hrnet/hrnet_config.py:46-73 (4 BASIC blocks each)
    (1, (48, 96)),
    (4, (48, 96, 192)),
    (3, (48, 96, 192, 384)),
)


def hrnet_ocr_param_shapes(n_classes=1):
    """OrderedDict name -> shape for `segmentation_model.*` when DETECTOR_TYPE = 'HRNet_OCR' (HRNet-W48 backbone +
    OCR head: hrnet_ocr/nets/hrnet.py:101-135, backbones/hrnet/hrnet_backbone.py:108-290,295-512,
    modules/spatial_ocr_block.py:114-280; every norm is nn.BatchNorm2d, bn_type 'torchbn')."""
    P = OrderedDict()

    def conv(name, co, ci, k, bias=False):
        P[name + ".weight"] = (co, ci, k, k)
        if bias:
            P[name + ".bias"] = (co,)

    b = "backbone."
    conv(b + "conv1", 64, 3, 3); _bn(P, b + "bn1", 64)
    conv(b + "conv2", 64, 64, 3); _bn(P, b + "bn2", 64)
    inpl = 64
    for i in range(4):                                   # layer1: 4 Bottlenecks, planes 64, expansion 4
        p = b + "layer1.%d" % i
        conv(p + ".conv1", 64, inpl, 1); _bn(P, p + ".bn1", 64)
        conv(p + ".conv2", 64, 64, 3); _bn(P, p + ".bn2", 64)
        conv(p + ".conv3", 256, 64, 1); _bn(P, p + ".bn3", 256)
        if i == 0:
            conv(p + ".downsample.0", 256, 64, 1); _bn(P, p + ".downsample.1", 256)
        inpl = 256
    pre = (256,)
    for si, (modules, chans) in enumerate(HRNET48_STAGES, 2):
        t = b + "transition%d" % (si - 1)
        for i, c in enumerate(chans):                    # _make_transition_layer :402-447
            if i < len(pre):
                if c != pre[i]:
                    conv(t + ".%d.0" % i, c, pre[i], 3); _bn(P, t + ".%d.1" % i, c)
            else:
                conv(t + ".%d.0.0" % i, c, pre[-1], 3); _bn(P, t + ".%d.0.1" % i, c)
        for m in range(modules):
            mp = b + "stage%d.%d" % (si, m)
            for bi, c in enumerate(chans):
                for k in range(4):
                    bp = mp + ".branches.%d.%d" % (bi, k)
                    conv(bp + ".conv1", c, c, 3); _bn(P, bp + ".bn1", c)
                    conv(bp + ".conv2", c, c, 3); _bn(P, bp + ".bn2", c)
            for i, ci in enumerate(chans):               # _make_fuse_layers :199-258
                for j, cj in enumerate(chans):
                    fp = mp + ".fuse_layers.%d.%d" % (i, j)
                    if j > i:
                        conv(fp + ".0", ci, cj, 1); _bn(P, fp + ".1", ci)
                    elif j < i:
                        for k in range(i - j):
                            co = ci if k == i - j - 1 else cj
                            conv(fp + ".%d.0" % k, co, cj, 3); _bn(P, fp + ".%d.1" % k, co)
        pre = chans
    conv("conv3x3.0", 512, 720, 3, bias=True); _bn(P, "conv3x3.1.0", 512)
    o = "ocr_distri_head.object_context_block."
    for name in ("f_pixel", "f_object"):
        conv(o + name + ".0", 256, 512, 1, bias=True); _bn(P, o + name + ".1.0", 256)
        conv(o + name + ".2", 256, 256, 1, bias=True); _bn(P, o + name + ".3.0", 256)
    conv(o + "f_down.0", 256, 512, 1, bias=True); _bn(P, o + "f_down.1.0", 256)
    conv(o + "f_up.0", 512, 256, 1, bias=True); _bn(P, o + "f_up.1.0", 512)
    conv("ocr_distri_head.conv_bn_dropout.0", 512, 1024, 1, bias=True); _bn(P, "ocr_distri_head.conv_bn_dropout.1.0", 512)
    conv("cls_head", n_classes, 512, 1, bias=True)
    conv("aux_head.0", 720, 720, 3, bias=True); _bn(P, "aux_head.1.0", 720)
    conv("aux_head.2", n_classes, 720, 1, bias=True)
    return P


_BUFFER_SUFFIXES = ("running_mean", "running_var", "num_batches_tracked")


class ParamTree(nn.Module):
    """Nested parameter holder reproducing dotted state_dict keys (never called as a network)."""

    def __init__(self, shapes=None):
        super().__init__()
        for name, shape in (shapes or {}).items():
            self._add(name.split("."), shape)

    def _add(self, parts, shape):
        if len(parts) == 1:
            leaf = parts[0]
            if leaf == "num_batches_tracked":
                self.register_buffer(leaf, torch.zeros((), dtype=torch.long))
            elif leaf in _BUFFER_SUFFIXES:
                self.register_buffer(leaf, torch.ones(shape) if leaf == "running_var" else torch.zeros(shape))
            else:
                self.register_parameter(leaf, nn.Parameter(torch.zeros(shape)))
            return
        child = self._modules.get(parts[0])
        if child is None:
            child = ParamTree()
            self.add_module(parts[0], child)
        child._add(parts[1:], shape)


def synth_state_dict(shapes, seed=1121, prefix=""):
    """Deterministic synthetic weights, independent of construction order (keyed by parameter name).

    Distributions mimic the reference's random init (kaiming-normal convs, kbpn.py:73-82,228-238;
    extractors.py:126-132) but also randomise biases / BN statistics / PReLU slopes so every parameter
    is exercised by the parity tests."""
    sd = OrderedDict()
    for name, shape in shapes.items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            t = torch.zeros((), dtype=torch.long)
        elif len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
            if shape[0] == 1:                                # 1-channel heads: widen the logit range
                t = t * 8.0
            if name in ("cls_head.weight", "aux_head.2.weight"):   # HRNet features are O(30) with random BN stacks:
                t = t * (0.03 / 8.0)                              # keep the head's logits O(1) instead of saturated
            if name.endswith("fe_cat.2.layer.weight"):       # keep the kernel refinement small (as after training):
                t = t * 0.01                                 # sum(pre + delta) stays near 1 -> well-conditioned /sum
        elif leaf == "running_var":
            t = 0.5 + torch.rand(shape, generator=g)
        elif leaf == "running_mean":
            t = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "weight" and shape == (1,):            # PReLU slope
            t = 0.01 + 0.24 * torch.rand(shape, generator=g)
        elif leaf == "weight":                               # BatchNorm gamma
            t = 0.3 + 0.5 * torch.rand(shape, generator=g)   # keeps the residual stack's activations O(1)
        elif leaf == "bias":
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            raise ValueError(name)
        sd[prefix + name] = t
    return sd
