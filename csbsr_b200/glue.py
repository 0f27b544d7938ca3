"""Differentiable glue ops of the training step on the csbsr_b200 kernels (csrc/glue.cu, csrc/support.cu): everything the
reference's train step runs between its convolutions (reference call sites in glue.cu's header).

torch.autograd is only the tape here: every node's forward and backward is one (or two) of our own launches on NHWC bf16 maps
[N, H, W, Cp] (Cp = channels padded to 64, padding channels zero) or fp32 NCHW images.  No aten kernel computes on activations.
"""
import ctypes as C

import torch

from . import _lib
from . import kernels as K
from .kernels import Fmap, round_up


def _call(name, *args):
    rc = getattr(_lib.lib(), name)(*args, _lib.stream_ptr())
    _lib.check(rc, name)
    _lib.count_launch(name)


def _chk(x):
    assert x.dtype == torch.bfloat16 and x.dim() == 4 and x.is_contiguous(), (x.dtype, x.shape, x.stride())
    return x


def _c(dy):
    return dy if dy.is_contiguous() else dy.contiguous()


_WS = {}


def _workspace(nbytes, device, tag):
    """Grow-only scratch buffer per (tag, device); kernels on one stream run in order, so reuse is safe."""
    key = (tag, str(device))
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = _WS[key] = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
    return ws


# ---------------------------------------------------------------------------------------------- elementwise
def _axpby(a, b, alpha, beta, relu=False):
    out = torch.empty_like(a)
    _call("csbsr_axpby", a.data_ptr(), b.data_ptr() if b is not None else None, out.data_ptr(), a.numel(), C.c_float(alpha),
          C.c_float(beta), int(relu))
    return out


class _AddFn(torch.autograd.Function):
    """out = [relu](a + beta * b)."""

    @staticmethod
    def forward(ctx, a, b, beta, relu):
        _chk(a), _chk(b)
        assert a.shape == b.shape
        out = _axpby(a, b, 1.0, beta, relu)
        ctx.beta, ctx.relu = beta, relu
        if relu:
            ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        if ctx.relu:
            (out,) = ctx.saved_tensors
            g = torch.empty_like(dy)
            _call("csbsr_act_bwd", dy.data_ptr(), out.data_ptr(), g.data_ptr(), dy.numel(), C.c_float(0.0))
            dy = g
        db = None
        if ctx.needs_input_grad[1]:
            db = dy if ctx.beta == 1.0 else _axpby(dy, None, ctx.beta, 0.0)
        return (dy if ctx.needs_input_grad[0] else None), db, None, None


def add(a, b, relu=False):
    return _AddFn.apply(a, b, 1.0, relu)


def sub(a, b):
    return _AddFn.apply(a, b, -1.0, False)


class _ActFn(torch.autograd.Function):
    """ReLU (slope 0) / LeakyReLU as a stand-alone node (HRNet fusion sums); the convs fuse theirs into the epilogue."""

    @staticmethod
    def forward(ctx, x, slope):
        _chk(x)
        if slope == 0.0:
            y = _axpby(x, None, 1.0, 0.0, relu=True)
        else:
            y = torch.empty_like(x)                       # leaky: y = x * (x > 0 ? 1 : slope) == act_bwd(dy = x, y = x)
            _call("csbsr_act_bwd", x.data_ptr(), x.data_ptr(), y.data_ptr(), x.numel(), C.c_float(slope))
        ctx.slope = slope
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = _c(dy)
        dx = torch.empty_like(dy)
        _call("csbsr_act_bwd", dy.data_ptr(), y.data_ptr(), dx.data_ptr(), dy.numel(), C.c_float(ctx.slope))
        return dx, None


def relu(x):
    return _ActFn.apply(x, 0.0)


def leaky_relu(x, slope):
    return _ActFn.apply(x, float(slope))


class _SftFn(torch.autograd.Function):
    """SFTlayer.forward (kbpn.py:516-518) / SFTLikeBlock (blocks.py:118-120): f * sigmoid(scale) + shift."""

    @staticmethod
    def forward(ctx, f, s, t):
        _chk(f), _chk(s), _chk(t)
        out = torch.empty_like(f)
        _call("csbsr_sft_combine", f.data_ptr(), s.data_ptr(), t.data_ptr(), out.data_ptr(), f.numel())
        ctx.save_for_backward(f, s)
        return out

    @staticmethod
    def backward(ctx, dy):
        f, s = ctx.saved_tensors
        dy = _c(dy)
        df, ds = torch.empty_like(f), torch.empty_like(s)
        _call("csbsr_sft_combine_bwd", dy.data_ptr(), f.data_ptr(), s.data_ptr(), df.data_ptr(), ds.data_ptr(), f.numel())
        return df, ds, dy


def sft_combine(f, scale_logits, shift):
    return _SftFn.apply(f, scale_logits, shift)


# ---------------------------------------------------------------------------------------------- concat / slice
def _window_copy(src, so, dst, do, c, zero_tail=0):
    rows = src.shape[0] * src.shape[1] * src.shape[2]
    _call("csbsr_window_copy", src.data_ptr(), src.shape[3], so, dst.data_ptr(), dst.shape[3], do, c, zero_tail, rows)


class _ConcatFn(torch.autograd.Function):
    """torch.cat along channels of NHWC maps, taking the first real[i] (multiple of 8) channels of part i; the result is
    zero-padded to a multiple of 64 channels.  Backward hands every part its slice (zero in its own padding channels)."""

    @staticmethod
    def forward(ctx, real, *parts):
        n, h, w, _ = parts[0].shape
        total = sum(real)
        cp = round_up(total, 64)
        out = torch.empty((n, h, w, cp), dtype=torch.bfloat16, device=parts[0].device)
        off = 0
        for i, (p, r) in enumerate(zip(parts, real)):
            _chk(p)
            assert r % 8 == 0 and r <= p.shape[3]
            _window_copy(p, 0, out, off, r, zero_tail=(cp - total) if i == len(parts) - 1 else 0)
            off += r
        ctx.real, ctx.pads = real, [p.shape[3] for p in parts]
        return out

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        n, h, w, _ = dy.shape
        grads, off = [], 0
        for i, (r, cp) in enumerate(zip(ctx.real, ctx.pads)):
            if ctx.needs_input_grad[1 + i]:
                g = torch.empty((n, h, w, cp), dtype=torch.bfloat16, device=dy.device)
                _window_copy(dy, off, g, 0, r, zero_tail=cp - r)
                grads.append(g)
            else:
                grads.append(None)
            off += r
        return (None, *grads)


def concat(parts, real=None):
    real = tuple(real) if real is not None else tuple(p.shape[3] for p in parts)
    return _ConcatFn.apply(real, *parts)


# ---------------------------------------------------------------------------------------------- resampling / pooling
class _BilinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, oh, ow, align):
        _chk(x)
        n, h, w, cp = x.shape
        y = torch.empty((n, oh, ow, cp), dtype=torch.bfloat16, device=x.device)
        K.bilinear(Fmap(x), Fmap(y), align_corners=align)
        ctx.cfg = (n, h, w, oh, ow, cp, align)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, h, w, oh, ow, cp, align = ctx.cfg
        dy = _c(dy)
        dx = torch.empty((n, h, w, cp), dtype=torch.bfloat16, device=dy.device)
        _call("csbsr_bilinear_nhwc_bwd", dy.data_ptr(), dx.data_ptr(), n, h, w, oh, ow, cp, cp, 0, cp, 0, int(align))
        return dx, None, None, None


def bilinear(x, size, align_corners=False):
    return _BilinearFn.apply(x, int(size[0]), int(size[1]), bool(align_corners))


class _AdaptivePoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, s):
        _chk(x)
        n, h, w, cp = x.shape
        y = torch.empty((n, s, s, cp), dtype=torch.bfloat16, device=x.device)
        K.adaptive_avgpool(Fmap(x), Fmap(y), s)
        ctx.cfg = (n, h, w, s, cp)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, h, w, s, cp = ctx.cfg
        dy = _c(dy)
        dx = torch.empty((n, h, w, cp), dtype=torch.bfloat16, device=dy.device)
        _call("csbsr_adaptive_avgpool_nhwc_bwd", dy.data_ptr(), dx.data_ptr(), n, h, w, s, cp, cp, 0, cp, 0)
        return dx, None


def adaptive_avgpool(x, s):
    return _AdaptivePoolFn.apply(x, int(s))


class _MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _chk(x)
        n, h, w, cp = x.shape
        y = torch.empty((n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, cp), dtype=torch.bfloat16, device=x.device)
        K.maxpool3s2(Fmap(x), Fmap(y))
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        n, h, w, cp = x.shape
        dy = _c(dy)
        dx = torch.empty_like(x)
        _call("csbsr_maxpool3s2_nhwc_bwd", x.data_ptr(), dy.data_ptr(), dx.data_ptr(), n, h, w, cp, cp, 0, cp, 0, cp, 0)
        return dx


def maxpool3s2(x):
    return _MaxPoolFn.apply(x)


class _GapFn(torch.autograd.Function):
    """nn.AdaptiveAvgPool2d(1) over the first c channels -> fp32 [N, c]."""

    @staticmethod
    def forward(ctx, x, c):
        _chk(x)
        out = torch.empty((x.shape[0], c), dtype=torch.float32, device=x.device)
        K.gap(Fmap(x), out, c)
        ctx.cfg = (tuple(x.shape), c)
        return out

    @staticmethod
    def backward(ctx, dy):
        (n, h, w, cp), c = ctx.cfg
        v = (dy.float() * (1.0 / (h * w))).contiguous()                 # [N, c]: tiny
        dx = torch.empty((n, h, w, cp), dtype=torch.bfloat16, device=dy.device)
        K.broadcast_vec(v, Fmap(dx))
        return dx, None


def gap(x, c):
    return _GapFn.apply(x, int(c))


# ---------------------------------------------------------------------------------------------- Dropout2d
class DropoutState:
    """Seed + DEVICE step counter of the Dropout2d masks: the counter is advanced by a kernel at the start of every training
    forward, so a CUDA-graph replay of the step draws new masks."""

    def __init__(self, device, seed=1121):
        self.seed = int(seed)
        self.counter = torch.zeros(1, dtype=torch.int64, device=device)
        self.salt = 0

    def begin_step(self):
        _call("csbsr_counter_inc", self.counter.data_ptr())
        self.salt = 0

    def next_salt(self):
        self.salt += 1
        return self.salt


class _Dropout2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, c, p, state):
        _chk(x)
        n, h, w, cp = x.shape
        scale = torch.empty((n, cp), dtype=torch.float32, device=x.device)
        _call("csbsr_dropout2d_mask", scale.data_ptr(), n, c, cp, C.c_float(p), state.seed, state.counter.data_ptr(), state.next_salt())
        y = torch.empty_like(x)
        _call("csbsr_channel_scale", x.data_ptr(), scale.data_ptr(), y.data_ptr(), n, h * w, cp)
        ctx.save_for_backward(scale)
        return y

    @staticmethod
    def backward(ctx, dy):
        (scale,) = ctx.saved_tensors
        dy = _c(dy)
        n, h, w, cp = dy.shape
        dx = torch.empty_like(dy)
        _call("csbsr_channel_scale", dy.data_ptr(), scale.data_ptr(), dx.data_ptr(), n, h * w, cp)
        return dx, None, None, None


def dropout2d(x, c, p, state):
    return _Dropout2dFn.apply(x, int(c), float(p), state)


# ---------------------------------------------------------------------------------------------- border classes
class _ExpandClassesFn(torch.autograd.Function):
    """[B, 2bw+1, 2bw+1, Cp] per-class responses -> [B, h, w, Cp]; backward = per-class sums of the gradient."""

    @staticmethod
    def forward(ctx, small, h, w, bw):
        _chk(small)
        n, k, _, cp = small.shape
        assert k == 2 * bw + 1
        out = torch.empty((n, h, w, cp), dtype=torch.bfloat16, device=small.device)
        _call("csbsr_expand_classes", small.data_ptr(), out.data_ptr(), n, h, w, bw, cp)
        ctx.cfg = (n, h, w, bw, cp)
        return out

    @staticmethod
    def backward(ctx, dy):
        n, h, w, bw, cp = ctx.cfg
        dy = _c(dy)
        need = _lib.lib().csbsr_expand_classes_workspace_bytes(n, h, bw, cp)
        ws = _workspace(need, dy.device, "expand_classes")
        ds = torch.empty((n, 2 * bw + 1, 2 * bw + 1, cp), dtype=torch.bfloat16, device=dy.device)
        _call("csbsr_expand_classes_bwd", dy.data_ptr(), ds.data_ptr(), n, h, w, bw, cp, ws.data_ptr(), need)
        return ds, None, None, None


def expand_classes(small, h, w, bw):
    return _ExpandClassesFn.apply(small, int(h), int(w), int(bw))


# ---------------------------------------------------------------------------------------------- layouts
class _ToNhwcFn(torch.autograd.Function):
    """fp32 NCHW -> NHWC bf16 with zero-padded channels."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous().float()
        n, c, h, w = x.shape
        y = torch.empty((n, h, w, round_up(c, 64)), dtype=torch.bfloat16, device=x.device)
        K.nchw_to_nhwc(x, Fmap(y))
        ctx.c = c
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        n, h, w, cp = dy.shape
        dx = torch.empty((n, ctx.c, h, w), dtype=torch.float32, device=dy.device)
        _call("csbsr_nhwc_bf16_to_nchw_f32", dy.data_ptr(), dx.data_ptr(), n, h * w, ctx.c, cp, 0)
        return dx


class _ToNchwFn(torch.autograd.Function):
    """NHWC bf16 (padded) -> fp32 NCHW with the first c channels."""

    @staticmethod
    def forward(ctx, x, c):
        _chk(x)
        n, h, w, cp = x.shape
        y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
        _call("csbsr_nhwc_bf16_to_nchw_f32", x.data_ptr(), y.data_ptr(), n, h * w, c, cp, 0)
        ctx.cp = cp
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous().float()
        n, c, h, w = dy.shape
        dx = torch.empty((n, h, w, ctx.cp), dtype=torch.bfloat16, device=dy.device)
        K.nchw_to_nhwc(dy, Fmap(dx))
        return dx, None


def to_nhwc(x):
    return _ToNhwcFn.apply(x)


def to_nchw(x, c):
    return _ToNchwFn.apply(x, int(c))


# ---------------------------------------------------------------------------------------------- instance norm
class _InstanceNormFn(torch.autograd.Function):
    """nn.InstanceNorm2d(3) of MetaSRModel.norm_sr (build_model.py:135-137): eps 1e-5, biased variance, no affine."""

    @staticmethod
    def forward(ctx, x, eps):
        x = x.contiguous()
        n, c, h, w = x.shape
        mean = torch.empty(n * c, dtype=torch.float32, device=x.device)
        rstd = torch.empty(n * c, dtype=torch.float32, device=x.device)
        K.clip_instnorm_stats(x, mean, rstd, do_clip=False, eps=eps)
        y = torch.empty_like(x)
        _call("csbsr_instnorm_apply", x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), y.data_ptr(), n * c, h * w)
        ctx.save_for_backward(x, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd = ctx.saved_tensors
        dy = dy.contiguous()
        n, c, h, w = x.shape
        need = _lib.lib().csbsr_instnorm_bwd_workspace_bytes(n * c)
        ws = _workspace(need, x.device, "instnorm")
        dx = torch.empty_like(x)
        _call("csbsr_instnorm_bwd", dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dx.data_ptr(), n * c, h * w,
              ws.data_ptr(), need)
        return dx, None


def instance_norm(x, eps=1e-5):
    return _InstanceNormFn.apply(x, float(eps))


# ---------------------------------------------------------------------------------------------- bias gradient
def bias_grad(dy, c):
    """sum over N*H*W of the first c channels of an NHWC bf16 map -> fp32 [c]."""
    n, h, w, cp = dy.shape
    need = _lib.lib().csbsr_colsum_workspace_bytes(c)
    ws = _workspace(need, dy.device, "colsum")
    out = torch.empty(c, dtype=torch.float32, device=dy.device)
    _call("csbsr_bias_grad", dy.data_ptr(), cp, 0, c, n * h * w, out.data_ptr(), ws.data_ptr(), need)
    return out
