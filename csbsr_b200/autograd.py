"""Differentiable conv / transposed-conv ops on the tcgen05 engine (training path).

Activations are NHWC bf16 tensors [N, H, W, Cp] with Cp = round_up(C, 64); the padding channels are always zero
(zero weight rows / bias entries produce them, ReLU / PReLU / add / mul keep them).  Parameters stay fp32 in the
reference's layouts (nn.Conv2d [Cout, Cin, R, S], nn.ConvTranspose2d [Cin, Cout, R, S]) so the optimizer and the
checkpoints are unchanged; they are packed to bf16 per call.

forward   y  = conv(x, w) + b                       -> csbsr_conv_igemm
backward  dx = conv^T(dy, w)                        -> csbsr_conv_igemm with the transposed / flipped packing
          dw = sum_pix dy (x) x                     -> csbsr_conv_wgrad
          db = sum_pix dy                           -> csbsr_bias_grad (fixed-order column sums)
An activation (ReLU / LeakyReLU) can ride in the forward epilogue; its backward masks dy by the sign of the saved output
(csbsr_act_bwd) before dgrad / wgrad / bias-grad.
Replaces autograd through cuDNN for nn.Conv2d / nn.ConvTranspose2d (reference model/modeling/kbpn.py:190-277,
pspnet_pytorch/extractors.py:37-70, trainer.py:57-72).
"""
import torch

import ctypes as C

from . import _lib
from . import kernels as K
from .kernels import Fmap, round_up

_ACT = {None: (K.ACT_NONE, 0.0), "relu": (K.ACT_RELU, 0.0)}


def _act_code(act):
    """act: None | 'relu' | ('lrelu', slope) -> (epilogue code, slope)"""
    if isinstance(act, tuple):
        return K.ACT_LEAKY, float(act[1])
    return _ACT[act]


def _mask_by_act(dy, y, slope):
    g = torch.empty_like(dy)
    _lib.check(_lib.lib().csbsr_act_bwd(dy.data_ptr(), y.data_ptr(), g.data_ptr(), dy.numel(), C.c_float(slope), _lib.stream_ptr()),
               "csbsr_act_bwd")
    _lib.count_launch("csbsr_act_bwd")
    return g


_SCRATCH = {}


def _scratch(nbytes, device, tag):
    """Grow-only scratch buffer per (tag, device) for the fixed-order reductions; launches on one stream run in order."""
    key = (tag, str(device))
    ws = _SCRATCH.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = _SCRATCH[key] = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
    return ws


def _bias_grad(dy, c):
    from .glue import bias_grad
    return bias_grad(dy, c)


def cpad(c):
    return round_up(c, 64)


def to_nhwc(x, dtype=torch.bfloat16):
    """fp32 NCHW -> NHWC bf16 with zero-padded channels (differentiable)."""
    n, c, h, w = x.shape
    y = x.permute(0, 2, 3, 1).to(dtype)
    if cpad(c) != c:
        y = torch.nn.functional.pad(y, (0, cpad(c) - c))
    return y.contiguous()


def to_nchw(x, c):
    """NHWC (padded) -> fp32 NCHW with the first `c` channels (differentiable)."""
    return x[..., :c].permute(0, 3, 1, 2).float()


def _out_size(h, k, stride, pad, dil):
    return (h + 2 * pad - dil * (k - 1) - 1) // stride + 1


class _Conv2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, dilation, act=None, cin_range=None):
        assert x.dtype == torch.bfloat16 and x.dim() == 4 and x.is_contiguous()
        co, ci, R, S = weight.shape
        if cin_range is not None:                    # only input channels [b0, b0 + ci) of the parameter take part
            ci = cin_range[1]
        n, h, w, cp = x.shape
        assert cp == cpad(ci), "input has %d channels, expected %d (padded %d)" % (cp, ci, cpad(ci))
        oh, ow = _out_size(h, R, stride, padding, dilation), _out_size(w, S, stride, padding, dilation)
        pc = K.pack_conv_train(weight, bias, stride=stride, padding=padding, dilation=dilation, cin_pad=cp, cout_pad=cpad(co),
                               cin_range=cin_range)
        y = Fmap.empty(n, oh, ow, cpad(co), device=x.device)
        code, slope = _act_code(act)
        K.conv(Fmap(x), pc, y, act=code, slope=slope)
        if act is None:
            ctx.save_for_backward(x, weight)
        else:
            ctx.save_for_backward(x, weight, y.t)
        ctx.cfg = (stride, padding, dilation, bias is not None, act is not None, slope, cin_range)
        return y.t

    @staticmethod
    def backward(ctx, dy):
        stride, padding, dilation, has_bias, has_act, slope, cin_range = ctx.cfg
        x, weight = ctx.saved_tensors[:2]
        co, ci, R, S = weight.shape
        if cin_range is not None:
            ci = cin_range[1]
        n, h, w, cp = x.shape
        dy = dy.contiguous()
        if has_act:
            dy = _mask_by_act(dy, ctx.saved_tensors[2], slope)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            if R == 8 and stride == 4 and padding == 2 and dilation == 1 and h == 4 * dy.shape[1] and w == 4 * dy.shape[2]:
                assert cin_range is None
                pc = K.pack_deconv8s4_train(weight, cin_pad=dy.shape[3], cout_pad=cp)  # conv 8/4/2 <-> convT 8/4/2
                g = Fmap.empty(n, h, w, cp, device=x.device)
                K.conv(Fmap(dy), pc, g)
            else:
                src = dy
                if stride > 1:                                                             # zero-stuffed gradient
                    hs, ws = h + 2 * padding - dilation * (R - 1), w + 2 * padding - dilation * (S - 1)
                    src = torch.zeros((n, hs, ws, dy.shape[3]), dtype=dy.dtype, device=dy.device)
                    src[:, ::stride, ::stride][:, :dy.shape[1], :dy.shape[2]] = dy
                pc = K.pack_conv_train(weight, None, stride=1, padding=dilation * (R - 1) - padding, dilation=dilation,
                                       cin_pad=dy.shape[3], cout_pad=cp, transpose_flip=True, cin_range=cin_range)
                g = Fmap.empty(n, h, w, cp, device=x.device)
                K.conv(Fmap(src), pc, g)
            dx = g.t
        if ctx.needs_input_grad[1]:
            taps = [(r * dilation - padding, s * dilation - padding) for r in range(R) for s in range(S)]
            tgt = K.wgrad_target(weight)
            if tgt is not None:                    # registered parameter: reduced straight into its .grad, in its own layout
                K.wgrad(Fmap(dy), Fmap(x), taps, stride=stride, into=(tgt, cin_range, 0))
            else:
                assert cin_range is None
                wg = K.wgrad(Fmap(dy), Fmap(x), taps, stride=stride)
                dw = wg[:co, :, :ci].reshape(co, R, S, ci).permute(0, 3, 1, 2).contiguous()
        if has_bias and ctx.needs_input_grad[2]:
            db = _bias_grad(dy, co)
        return dx, dw, db, None, None, None, None, None


class _Deconv8s4Fn(torch.autograd.Function):
    """nn.ConvTranspose2d(kernel 8, stride 4, padding 2): reference DeconvBlock (kbpn.py:273-277)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        assert x.dtype == torch.bfloat16 and x.is_contiguous()
        ci, co, R, S = weight.shape
        assert R == 8 and S == 8
        n, h, w, cp = x.shape
        assert cp == cpad(ci)
        pc = K.pack_deconv8s4_train(weight, bias, cin_pad=cp, cout_pad=cpad(co))
        y = Fmap.empty(n, 4 * h, 4 * w, cpad(co), device=x.device)
        K.conv(Fmap(x), pc, y)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return y.t

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        ci, co, R, S = weight.shape
        n, h, w, cp = x.shape
        dy = dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            pc = K.pack_conv_train(weight, None, stride=4, padding=2, cin_pad=dy.shape[3], cout_pad=cp)
            g = Fmap.empty(n, h, w, cp, device=x.device)
            K.conv(Fmap(dy), pc, g)
            dx = g.t
        if ctx.needs_input_grad[1]:
            taps = [(r - 2, s - 2) for r in range(8) for s in range(8)]
            tgt = K.wgrad_target(weight)
            if tgt is not None:
                K.wgrad(Fmap(x), Fmap(dy), taps, stride=4, into=(tgt, None, 0))
            else:
                wg = K.wgrad(Fmap(x), Fmap(dy), taps, stride=4)
                dw = wg[:ci, :, :co].reshape(ci, 8, 8, co).permute(0, 3, 1, 2).contiguous()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _bias_grad(dy, co)
        return dx, dw, db


class _TapExpConv3x3Fn(torch.autograd.Function):
    """3x3 / padding-1 conv with <= 4 output channels and no bias (KBlock.sr_reconst kbpn.py:361, output_conv :68): N = 3 would
    pay the whole 9-tap A-operand stream for three useful columns, forward, dgrad and wgrad alike.  Instead z = x . W_exp is ONE
    1x1 GEMM with 9 * 4 outputs (output t*4 + m = tap t of channel m at the same pixel), y gathers the nine shifted taps
    (csbsr_tapexp_gather_nhwc); backward scatters dy into the same layout, so dx = dz . W_exp^T and dW = dz^T x are 1x1 GEMMs."""
    CP = 4

    @staticmethod
    def forward(ctx, x, weight):
        assert x.dtype == torch.bfloat16 and x.dim() == 4 and x.is_contiguous()
        co, ci, R, S = weight.shape
        n, h, w, cp_in = x.shape
        assert (R, S) == (3, 3) and co <= _TapExpConv3x3Fn.CP and cp_in == cpad(ci)
        z = Fmap.empty(n, h, w, 64, device=x.device)
        K.conv(Fmap(x), K.pack_tapexp_train(weight, cp_in, _TapExpConv3x3Fn.CP), z)
        y = torch.empty((n, h, w, 64), dtype=torch.bfloat16, device=x.device)
        _lib.check(_lib.lib().csbsr_tapexp_gather_nhwc(z.ptr(), 64, y.data_ptr(), 64, n, h, w, _TapExpConv3x3Fn.CP, co,
                                                       _lib.stream_ptr()), "csbsr_tapexp_gather_nhwc")
        _lib.count_launch("csbsr_tapexp_gather_nhwc")
        ctx.save_for_backward(x, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        co, ci = weight.shape[:2]
        n, h, w, cp_in = x.shape
        cp = _TapExpConv3x3Fn.CP
        dy = dy.contiguous()
        dz = torch.empty((n, h, w, 64), dtype=torch.bfloat16, device=x.device)
        _lib.check(_lib.lib().csbsr_tapexp_scatter_nhwc(dy.data_ptr(), dy.shape[3], dz.data_ptr(), 64, n, h, w, cp, co,
                                                        _lib.stream_ptr()), "csbsr_tapexp_scatter_nhwc")
        _lib.count_launch("csbsr_tapexp_scatter_nhwc")
        dx = dw = None
        if ctx.needs_input_grad[0]:
            g = Fmap.empty(n, h, w, cp_in, device=x.device)
            K.conv(Fmap(dz), K.pack_tapexp_train(weight, cp_in, cp, transpose=True), g)
            dx = g.t
        if ctx.needs_input_grad[1]:
            tgt = K.wgrad_target(weight)
            if tgt is not None:
                K.wgrad(Fmap(dz), Fmap(x), [(0, 0)], into=(tgt, None, cp))
            else:
                wg = K.wgrad(Fmap(dz), Fmap(x), [(0, 0)])                   # [128][1][cp_in]
                dw = wg[:9 * cp, 0, :ci].reshape(9, cp, ci)[:, :co].permute(1, 2, 0).reshape(co, ci, 3, 3).contiguous()
        return dx, dw


def conv3x3_few_outputs(x, weight):
    return _TapExpConv3x3Fn.apply(x, weight)


def conv2d(x, weight, bias=None, stride=1, padding=0, dilation=1, act=None, cin_range=None):
    """act: None | 'relu' | ('lrelu', slope) -- applied in the conv epilogue.  cin_range = (b0, b): only input channels
    [b0, b0 + b) of `weight` (a parameter registered with kernels.register_params) are used -- the two halves of one conv."""
    if cin_range is not None and K.registered(weight) is None:
        weight, cin_range = weight[:, cin_range[0]:cin_range[0] + cin_range[1]].contiguous(), None
    return _Conv2dFn.apply(x, weight, bias, stride, padding, dilation, act, cin_range)


def deconv8s4(x, weight, bias=None):
    return _Deconv8s4Fn.apply(x, weight, bias)


class _PReLUFn(torch.autograd.Function):
    """Single-parameter nn.PReLU on a bf16 map; the slope gradient sum_{x<0} dy*x is accumulated in fp32 on the device
    (a bf16 reduction of that heavily cancelling sum is off by tens of percent)."""

    @staticmethod
    def forward(ctx, x, slope):
        from . import _lib
        assert x.dtype == torch.bfloat16 and x.is_contiguous() and slope.numel() == 1 and slope.dtype == torch.float32
        y = torch.empty_like(x)
        _lib.check(_lib.lib().csbsr_prelu_fwd(x.data_ptr(), y.data_ptr(), slope.data_ptr(), x.numel(), _lib.stream_ptr()),
                   "csbsr_prelu_fwd")
        _lib.count_launch("csbsr_prelu_fwd")
        ctx.save_for_backward(x, slope)
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import _lib
        x, slope = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        ds = torch.empty(1, dtype=torch.float32, device=x.device)
        ws = _scratch(_lib.lib().csbsr_prelu_bwd_workspace_bytes(), x.device, "prelu")
        _lib.check(_lib.lib().csbsr_prelu_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), slope.data_ptr(), ds.data_ptr(),
                                              x.numel(), ws.data_ptr(), _lib.stream_ptr()), "csbsr_prelu_bwd")
        _lib.count_launch("csbsr_prelu_bwd")
        return dx, ds.view(slope.shape)


def prelu(x, slope):
    return _PReLUFn.apply(x, slope)


class _BlurFn(torch.autograd.Function):
    """Per-sample depthwise blur, stride s, zero padding (k-1)/2: every image of the batch with its own k x k kernel
    (KBlock pseudo-LR, kbpn.py:395-402; Get_pseudo_lr of KBPNLoss, sr_loss_functions.py:73-102).  img fp32 [B,C,H,W],
    kvec fp32 [B,k*k] -> fp32 [B,C,ceil(H/s),ceil(W/s)]; gradients w.r.t. both on the device kernels of csrc/train.cu."""

    @staticmethod
    def forward(ctx, img, kvec, ksize, stride):
        from . import kernels as K
        img, kvec = img.contiguous(), kvec.contiguous()
        b, c, h, w = img.shape
        out = torch.empty((b, c, (h - 1) // stride + 1, (w - 1) // stride + 1), dtype=torch.float32, device=img.device)
        K.blur_per_sample(img, kvec, None, out, ksize, stride)
        ctx.save_for_backward(img, kvec)
        ctx.cfg = (ksize, stride)
        return out

    @staticmethod
    def backward(ctx, dy):
        from . import _lib
        img, kvec = ctx.saved_tensors
        ksize, stride = ctx.cfg
        b, c, h, w = img.shape
        dy = dy.contiguous().float()
        dx = dk = None
        L = _lib.lib()
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(img)
            _lib.check(L.csbsr_blur_ps_bwd_input(dy.data_ptr(), kvec.data_ptr(), dx.data_ptr(), b, c, h, w, ksize, stride,
                                                 _lib.stream_ptr()), "csbsr_blur_ps_bwd_input")
            _lib.count_launch("csbsr_blur_ps_bwd_input")
        if ctx.needs_input_grad[1]:
            dk = torch.empty_like(kvec)
            ws = _scratch(L.csbsr_blur_ps_bwd_kernel_workspace_bytes(b, c, h, w, ksize, stride), img.device, "blur_dk")
            _lib.check(L.csbsr_blur_ps_bwd_kernel(img.data_ptr(), dy.data_ptr(), dk.data_ptr(), b, c, h, w, ksize, stride,
                                                  ws.data_ptr(), _lib.stream_ptr()), "csbsr_blur_ps_bwd_kernel")
            _lib.count_launch("csbsr_blur_ps_bwd_kernel")
        return dx, dk, None, None


class _ResizeAAFn(torch.autograd.Function):
    """FactorResize(factor, 'bicubic') = antialiased bicubic downscale (transforms.py:516-531) with its transpose as backward."""

    @staticmethod
    def forward(ctx, x, factor):
        from . import _lib
        x = x.contiguous()
        n, c, h, w = x.shape
        oh, ow = int(h / factor), int(w / factor)
        out = torch.empty((n, c, oh, ow), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().csbsr_resize_bicubic_aa(x.data_ptr(), out.data_ptr(), n * c, h, w, oh, ow, 0, _lib.stream_ptr()),
                   "csbsr_resize_bicubic_aa")
        _lib.count_launch("csbsr_resize_bicubic_aa")
        ctx.shape = (n, c, h, w, oh, ow)
        return out

    @staticmethod
    def backward(ctx, dy):
        from . import _lib
        n, c, h, w, oh, ow = ctx.shape
        dy = dy.contiguous().float()
        dx = torch.empty((n, c, h, w), dtype=torch.float32, device=dy.device)
        _lib.check(_lib.lib().csbsr_resize_bicubic_aa_bwd(dy.data_ptr(), dx.data_ptr(), n * c, h, w, oh, ow, _lib.stream_ptr()),
                   "csbsr_resize_bicubic_aa_bwd")
        _lib.count_launch("csbsr_resize_bicubic_aa_bwd")
        return dx, None


def blur_per_sample(img, kvec, ksize, stride):
    return _BlurFn.apply(img, kvec, ksize, stride)


def resize_aa(x, factor):
    return _ResizeAAFn.apply(x, factor)


class _BatchNormFn(torch.autograd.Function):
    """nn.BatchNorm2d on an NHWC bf16 map with optional fused residual add and ReLU:  y = relu?(bn(x) + res?).
    training=True: batch statistics (and the in-place momentum update of the running buffers); False: running statistics."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, res, training, momentum, eps, relu):
        from . import _lib
        assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.dim() == 4
        n, h, w, pitch = x.shape
        c, m = gamma.numel(), n * h * w
        L = _lib.lib()
        if training:
            mean = torch.empty(c, dtype=torch.float32, device=x.device)
            rstd = torch.empty(c, dtype=torch.float32, device=x.device)
            ws = _scratch(L.csbsr_bn_workspace_bytes(c), x.device, "bn")
            _lib.check(L.csbsr_bn_stats(x.data_ptr(), pitch, c, m, eps, momentum, mean.data_ptr(), rstd.data_ptr(),
                                        running_mean.data_ptr() if running_mean is not None else None,
                                        running_var.data_ptr() if running_var is not None else None, ws.data_ptr(),
                                        _lib.stream_ptr()), "csbsr_bn_stats")
            _lib.count_launch("csbsr_bn_stats")
        else:
            mean = running_mean.detach().float().contiguous()
            rstd = torch.rsqrt(running_var.detach().float() + eps).contiguous()
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        if res is not None:
            res = res.contiguous()
            assert res.shape == x.shape and res.dtype == torch.bfloat16
        y = torch.empty_like(x)
        _lib.check(L.csbsr_bn_apply(x.data_ptr(), res.data_ptr() if res is not None else None, y.data_ptr(), mean.data_ptr(),
                                    rstd.data_ptr(), g32.data_ptr(), b32.data_ptr(), pitch, c, m, int(relu), _lib.stream_ptr()),
                   "csbsr_bn_apply")
        _lib.count_launch("csbsr_bn_apply")
        ctx.save_for_backward(x, y if relu else None, g32, mean, rstd)
        ctx.cfg = (bool(training), res is not None, c)
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import _lib
        x, y, g32, mean, rstd = ctx.saved_tensors
        training, has_res, c = ctx.cfg
        dy = dy.contiguous()
        n, h, w, pitch = x.shape
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if has_res else None
        dgamma = torch.empty(c, dtype=torch.float32, device=x.device)
        dbeta = torch.empty(c, dtype=torch.float32, device=x.device)
        ws = _scratch(_lib.lib().csbsr_bn_workspace_bytes(c), x.device, "bn")
        _lib.check(_lib.lib().csbsr_bn_backward(dy.data_ptr(), x.data_ptr(), y.data_ptr() if y is not None else None,
                                                mean.data_ptr(), rstd.data_ptr(), g32.data_ptr(), pitch, c, n * h * w,
                                                int(training), dx.data_ptr(), dres.data_ptr() if has_res else None,
                                                dgamma.data_ptr(), dbeta.data_ptr(), ws.data_ptr(), _lib.stream_ptr()),
                   "csbsr_bn_backward")
        _lib.count_launch("csbsr_bn_backward")
        return dx, dgamma, dbeta, None, None, dres, None, None, None, None


def batch_norm(x, gamma, beta, running_mean, running_var, training, momentum=0.1, eps=1e-5, relu=False, res=None):
    return _BatchNormFn.apply(x, gamma, beta, running_mean, running_var, res, training, momentum, eps, relu)
