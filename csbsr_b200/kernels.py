"""Python launchers for the C-ABI kernels: weight packing + descriptor filling. No math happens here.

Activations live in HBM as NHWC bf16 "feature maps" (`Fmap`): a [N,H,W,P] tensor plus a channel
window, so that the reference's torch.cat calls (kbpn.py:173-186) become writes into channel slices.
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import (ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, OUT_BF16_NHWC, OUT_F32_NCHW, OUT_F32_NHWC,
                   ConvDesc)


PROFILE = None
PROFILE_WG = None          # list of (label, flops, start event, end event) per csbsr_conv_wgrad call when set        # set to a list to collect (label, flops, start_event, end_event) per conv launch


def round_up(v, m):
    return (v + m - 1) // m * m


class Fmap:
    """A channel window [coff, coff+c) of an NHWC bf16 tensor."""

    def __init__(self, t, coff=0, c=None):
        assert t.dim() == 4 and t.dtype == torch.bfloat16 and t.is_contiguous()
        self.t = t
        self.coff = coff
        self.c = t.shape[3] - coff if c is None else c
        assert 0 <= coff and coff + self.c <= t.shape[3]

    @staticmethod
    def empty(n, h, w, c, device="cuda", zero=False):
        f = torch.zeros if zero else torch.empty
        return Fmap(f((n, h, w, c), dtype=torch.bfloat16, device=device))

    @property
    def n(self):
        return self.t.shape[0]

    @property
    def h(self):
        return self.t.shape[1]

    @property
    def w(self):
        return self.t.shape[2]

    @property
    def pitch(self):
        return self.t.shape[3]

    def window(self, coff, c):
        return Fmap(self.t, self.coff + coff, c)

    def ptr(self):
        return self.t.data_ptr()

    def to_nchw_f32(self, c=None):
        c = self.c if c is None else c
        return self.t[..., self.coff:self.coff + c].permute(0, 3, 1, 2).float().contiguous()

    @staticmethod
    def from_nchw(x, cpad=None):
        n, c, h, w = x.shape
        cpad = cpad or round_up(c, 64)
        t = torch.zeros((n, h, w, cpad), dtype=torch.bfloat16, device=x.device)
        t[..., :c] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
        return Fmap(t, 0, cpad)


class PlanarWin:
    """Channel window [coff, coff+c) of a planar fp32 [N,P,H,W] tensor (conv output / fp32 residual)."""

    def __init__(self, t, coff=0, c=None):
        assert t.dim() == 4 and t.dtype == torch.float32 and t.is_contiguous()
        self.t, self.coff = t, coff
        self.c = t.shape[1] - coff if c is None else c
        assert 0 <= coff and coff + self.c <= t.shape[1]


class F32Map:
    """fp32 NHWC conv output [N,H,W,P] (used for the per-sample border-class biases)."""

    def __init__(self, t):
        assert t.dim() == 4 and t.dtype == torch.float32 and t.is_contiguous()
        self.t = t


class PackedConv:
    """Weights of one conv / transposed conv packed K-major per tap: [taps][cout_pad][cin_pad] bf16."""

    def __init__(self, wp, taps, nphases, ntaps, stride, os, ooh, oow, cout, bias=None, macs_per_pixel=None):
        self.wp = wp                      # [w_taps, cout_pad, cin_pad] bf16, contiguous
        self.taps = taps                  # list of (dh, dw, widx), length nphases*ntaps
        self.nphases, self.ntaps = nphases, ntaps
        self.stride, self.os = stride, os
        self.ooh, self.oow = ooh, oow
        self.cout = cout
        self.bias = bias                  # fp32 [cout_pad] or None
        # useful (unpadded, non-zero-weight) multiply-accumulates per output pixel of ONE phase
        self.macs_per_pixel = macs_per_pixel if macs_per_pixel is not None else ntaps * wp.shape[2] * cout

    @property
    def cout_pad(self):
        return self.wp.shape[1]

    @property
    def cin_pad(self):
        return self.wp.shape[2]


def to_device(obj, device):
    """Move every tensor of a packed-weight tree (dicts / lists / tuples / PackedConv) to `device`.  The engines pack on
    the host (weight folding, padding, transposition are one-off host work) and ship the results with plain H2D copies, so
    loading a checkpoint launches no device kernels."""
    if isinstance(obj, torch.Tensor):
        return obj.to(device)
    if isinstance(obj, PackedConv):
        obj.wp = obj.wp.to(device)
        if obj.bias is not None:
            obj.bias = obj.bias.to(device)
        return obj
    if isinstance(obj, dict):
        return {k: to_device(v, device) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(to_device(v, device) for v in obj)
    return obj


def _pad_bias(bias, cout_pad, device):
    if bias is None:
        return None
    b = torch.zeros(cout_pad, dtype=torch.float32, device=device)
    b[:bias.numel()] = bias.detach().float()
    return b


def pack_conv(weight, bias=None, stride=1, padding=0, dilation=1, cin_pad=None, cout_pad=None, scale=None,
              cin_map=None):
    """nn.Conv2d weight [Cout,Cin,R,S] -> PackedConv. `scale` (per-cout) folds an eval-mode BatchNorm.
    `cin_map`: optional list of (src_start, src_len, dst_start) placing input-channel ranges in the padded K axis."""
    w = weight.detach().float()
    if scale is not None:
        w = w * scale.view(-1, 1, 1, 1)
    cout, cin, R, S = w.shape
    cout_pad = cout_pad or round_up(cout, 16)
    if cin_map is None:
        cin_map = [(0, cin, 0)]
    need = max(d + l for _, l, d in cin_map)
    cin_pad = cin_pad or round_up(need, 64)
    wp = torch.zeros((R * S, cout_pad, cin_pad), dtype=torch.float32, device=w.device)
    wt = w.permute(2, 3, 0, 1).reshape(R * S, cout, cin)
    for s0, ln, d0 in cin_map:
        wp[:, :cout, d0:d0 + ln] = wt[:, :, s0:s0 + ln]
    taps = [(r * dilation - padding, s * dilation - padding, r * S + s) for r in range(R) for s in range(S)]
    return PackedConv(wp.to(torch.bfloat16).contiguous(), taps, 1, R * S, stride, 1, [0], [0], cout,
                      _pad_bias(bias, cout_pad, w.device), macs_per_pixel=R * S * cin * cout)


def pack_deconv8s4(weight, bias=None, cin_pad=None, cout_pad=None):
    """nn.ConvTranspose2d(k=8, s=4, p=2) weight [Cin,Cout,8,8] -> 16 output phases of a 2x2 conv.

    out[4q+rho] = sum_{ih} x[ih] * w[.., r = 4q+rho+2-4ih]; for rho in {0,1}: ih in {q-1 (r=rho+6), q (r=rho+2)},
    for rho in {2,3}: ih in {q (r=rho+2), q+1 (r=rho-2)}  (kbpn.py:274-277 DeconvBlock, 8/4/2 setting :23-26)."""
    w = weight.detach().float()
    cin, cout, R, S = w.shape
    assert R == 8 and S == 8
    cout_pad = cout_pad or round_up(cout, 16)
    cin_pad = cin_pad or round_up(cin, 64)
    wp = torch.zeros((64, cout_pad, cin_pad), dtype=torch.float32, device=w.device)
    wp[:, :cout, :cin] = w.permute(2, 3, 1, 0).reshape(64, cout, cin)

    def axis_taps(rho):
        return [(-1, rho + 6), (0, rho + 2)] if rho < 2 else [(0, rho + 2), (1, rho - 2)]

    taps, ooh, oow = [], [], []
    for rh in range(4):
        for rw in range(4):
            for dh, r in axis_taps(rh):
                for dw, s in axis_taps(rw):
                    taps.append((dh, dw, r * 8 + s))
            ooh.append(rh)
            oow.append(rw)
    return PackedConv(wp.to(torch.bfloat16).contiguous(), taps, 16, 4, 1, 4, ooh, oow, cout,
                      _pad_bias(bias, cout_pad, w.device), macs_per_pixel=4 * cin * cout)


def conv(x, pc, y, oh=None, ow=None, bias=None, bias_sn=0, bias_sc=0, cls_bw=0, act=ACT_NONE, slope=0.0,
         r0=None, rm=None, r1=None, r1_sign=1.0, r32=None, cout_store=None, block_n=0):
    """Launch the tcgen05 implicit-GEMM conv. `y` is an Fmap (bf16 NHWC) or an fp32 [N,C,H,W] tensor."""
    d = ConvDesc()
    d.x = x.ptr()
    d.n, d.h, d.w = x.n, x.h, x.w
    d.x_pitch, d.x_coff, d.cin = x.pitch, x.coff, pc.cin_pad
    assert x.c >= pc.cin_pad or x.coff + pc.cin_pad <= x.pitch, "input window narrower than packed K"
    assert pc.cin_pad % 64 == 0 or pc.cin_pad in (16, 32), "packed K must be a multiple of 64, or 32 / 16"
    d.wgt = pc.wp.data_ptr()
    d.w_taps, d.cout_pad = pc.wp.shape[0], pc.cout_pad
    d.nphases, d.ntaps, d.stride = pc.nphases, pc.ntaps, pc.stride
    d.os = pc.os
    merged = _deconv_merge_ok(x, pc, y, rm, r0, r1, oh, ow)
    if merged:
        # 8x8 / stride-4 transposed conv as 4 tap classes x 4 sub-phases: the sub-phases of a class share their 2x2 input
        # taps, so two of them form one N = 256 tile on one A operand (csbsr_conv_desc.nsub)
        taps_c, widx_c, ooh_c, oow_c = _DECONV_MERGED
        d.nphases, d.nsub = 4, 4
        for i, (dh, dw) in enumerate(taps_c):
            d.dh[i], d.dw[i] = dh, dw
        for i, wi in enumerate(widx_c):
            d.widx[i] = wi
        for i in range(16):
            d.ooh[i], d.oow[i] = ooh_c[i], oow_c[i]
    else:
        for i, (dh, dw, wi) in enumerate(pc.taps):
            d.dh[i], d.dw[i], d.widx[i] = dh, dw, wi
        for i in range(pc.nphases):
            d.ooh[i], d.oow[i] = pc.ooh[i], pc.oow[i]
    if isinstance(y, Fmap):
        d.out_mode = OUT_BF16_NHWC
        d.yh, d.yw = y.h, y.w
        d.y = y.ptr()
        d.y_pitch, d.y_coff = y.pitch, y.coff
        d.cout_store = cout_store if cout_store is not None else min(y.c, pc.cout_pad)
        assert y.n == x.n
    elif isinstance(y, F32Map):
        d.out_mode = OUT_F32_NHWC
        d.yh, d.yw = y.t.shape[1], y.t.shape[2]
        d.y = y.t.data_ptr()
        d.y_pitch, d.y_coff = y.t.shape[3], 0
        d.cout_store = cout_store if cout_store is not None else min(y.t.shape[3], pc.cout_pad)
        assert y.t.shape[0] == x.n
    elif isinstance(y, PlanarWin):
        d.out_mode = OUT_F32_NCHW
        d.yh, d.yw = y.t.shape[2], y.t.shape[3]
        d.y = y.t.data_ptr()
        d.y_pitch, d.y_coff, d.cout_store = y.t.shape[1], y.coff, y.c
        assert y.t.shape[0] == x.n
    else:
        assert y.dtype == torch.float32 and y.is_contiguous() and y.dim() == 4
        d.out_mode = OUT_F32_NCHW
        d.yh, d.yw = y.shape[2], y.shape[3]
        d.y = y.data_ptr()
        d.cout_store = y.shape[1]
        assert y.shape[0] == x.n
    if oh is None:
        oh = (d.yh - pc.ooh[0] + pc.os - 1) // pc.os if pc.os > 1 else d.yh
        ow = (d.yw - pc.oow[0] + pc.os - 1) // pc.os if pc.os > 1 else d.yw
    d.oh, d.ow = oh, ow
    b = bias if bias is not None else pc.bias
    if b is not None:
        assert b.dtype == torch.float32 and b.is_contiguous()
        d.bias = b.data_ptr()
        d.bias_sn, d.bias_sc, d.cls_bw = bias_sn, bias_sc, cls_bw
    d.act, d.slope = act, float(slope)
    keep = [b]
    for name, r in (("r0", r0), ("rm", rm), ("r1", r1)):
        if r is not None:
            assert r.n == x.n and r.h == d.yh and r.w == d.yw
            setattr(d, name, r.ptr())
            setattr(d, name + "_pitch", r.pitch)
            setattr(d, name + "_coff", r.coff)
    d.r1_sign = float(r1_sign)
    if r32 is not None:
        if isinstance(r32, PlanarWin):
            assert r32.c == d.cout_store and r32.t.shape[2:] == (d.yh, d.yw)
            d.r32 = r32.t.data_ptr()
            d.r32_pitch, d.r32_coff = r32.t.shape[1], r32.coff
        else:
            assert r32.dtype == torch.float32 and r32.is_contiguous()
            d.r32 = r32.data_ptr()
    d.block_n = block_n
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = _lib.lib().csbsr_conv_igemm(C.byref(d), _lib.stream_ptr())
    _lib.check(rc, "csbsr_conv_igemm")
    _lib.count_launch("csbsr_conv_igemm")
    if PROFILE is not None:
        e1.record()
        flops = 2.0 * x.n * d.oh * d.ow * pc.nphases * pc.ntaps * pc.cin_pad * pc.cout_pad
        useful = 2.0 * x.n * d.oh * d.ow * pc.nphases * pc.macs_per_pixel
        # algorithmic HBM bytes: the input window once, the stored output once, each residual operand once
        out_b = x.n * d.oh * d.ow * pc.nphases * d.cout_store * (2 if d.out_mode == OUT_BF16_NHWC else 4)
        nres = sum(r is not None for r in (r0, rm, r1)) + (2 if r32 is not None else 0)
        nbytes = x.n * x.h * x.w * pc.cin_pad * 2.0 + out_b * (1 + nres)
        PROFILE.append(("conv n%d %dx%d cin%d cout%d taps%dx%d s%d" % (x.n, d.oh, d.ow, pc.cin_pad, pc.cout_pad,
                                                                        pc.nphases, pc.ntaps, pc.stride), flops, e0, e1,
                        useful, nbytes))
    return y


# ------------------------------------------------------------------ training-step weight cache
# Parameters registered by the optimizer (engine/optim.py::FusedAdam) keep their packed bf16 GEMM operands across the step:
# every (parameter, layout) pair is packed on first use and afterwards refreshed by ONE multi-tensor launch at the first
# request that follows an optimizer step (invalidate_packed()), instead of one launch per use (297 per step).
_REG = {}                      # data_ptr -> parameter
_PACKS = {}                    # (data_ptr, A, Btot, b0, B, R, S, rows_pad, cols_pad, mode) -> entry
_PACK_STATE = {"epoch": 0, "dev": None, "n": 0, "total": 0, "taps": 1, "dirty": True}


def register_params(params):
    """Called by the optimizer that owns the flat parameter / gradient buffers; a new optimizer replaces the previous set."""
    _REG.clear()
    _PACKS.clear()
    _PACK_STATE.update(dev=None, n=0, total=0, taps=1, dirty=True)
    for p in params:
        _REG[p.data_ptr()] = p


def registered(weight):
    p = _REG.get(weight.data_ptr())
    return p if (p is not None and p.shape == weight.shape) else None


def invalidate_packed():
    """The parameters changed (optimizer step / load_state_dict): cached packs are stale from now on."""
    _PACK_STATE["epoch"] += 1


def flush_pack_table():
    """Upload the job table of the multi-tensor pack if entries were added (call before a CUDA-graph capture)."""
    st = _PACK_STATE
    if not st["dirty"] or not _PACKS:
        return
    L = _lib.lib()
    jb = L.csbsr_pack_job_bytes()
    buf = (C.c_ubyte * (jb * len(_PACKS)))()
    start, taps = 0, 1
    for i, (key, e) in enumerate(_PACKS.items()):
        _, a, btot, b0, b, R, S, rows_pad, cols_pad, mode = key
        _lib.check(L.csbsr_pack_job_fill(C.byref(buf, i * jb), e["w"].data_ptr(), e["out"].data_ptr(), a, b, btot, b0, R, S,
                                         rows_pad, cols_pad, mode, start), "csbsr_pack_job_fill")
        start += L.csbsr_pack_job_tiles(a, b, R, S, rows_pad, cols_pad, mode)
        taps = max(taps, R * S)
    dev = next(iter(_PACKS.values()))["out"].device
    st["dev"] = torch.frombuffer(buf, dtype=torch.uint8).clone().to(dev)
    st["n"], st["total"], st["taps"], st["dirty"] = len(_PACKS), start, taps, False


def _refresh_all():
    st = _PACK_STATE
    flush_pack_table()
    _call("csbsr_pack_weights_multi", st["dev"].data_ptr(), st["n"], st["total"], st["taps"])
    for e in _PACKS.values():
        e["epoch"], e["version"] = st["epoch"], e["w"]._version


def _pack_device(weight, rows_pad, cols_pad, mode, cin_range=None):
    """Packing of a contiguous fp32 CUDA parameter (csbsr_pack_weights*); cin_range = (b0, b) packs a window of its second axis.
    mode & 7 >= 3: tap-expanded 3x3 weight (one packed "tap", cp = mode >> 3)."""
    a, btot, R, S = weight.shape
    if (mode & 7) >= 3:
        assert (R, S) == (3, 3)
        R = S = 1
    b0, b = cin_range if cin_range is not None else (0, btot)
    reg = registered(weight)
    if reg is None:
        out = torch.empty((R * S, rows_pad, cols_pad), dtype=torch.bfloat16, device=weight.device)
        _call("csbsr_pack_weights_window", weight.data_ptr(), out.data_ptr(), a, b, btot, b0, R, S, rows_pad, cols_pad, mode)
        return out
    key = (weight.data_ptr(), a, btot, b0, b, R, S, rows_pad, cols_pad, mode)
    e = _PACKS.get(key)
    if e is None:
        out = torch.empty((R * S, rows_pad, cols_pad), dtype=torch.bfloat16, device=weight.device)
        _call("csbsr_pack_weights_window", weight.data_ptr(), out.data_ptr(), a, b, btot, b0, R, S, rows_pad, cols_pad, mode)
        _PACKS[key] = {"w": reg, "out": out, "epoch": _PACK_STATE["epoch"], "version": reg._version}
        _PACK_STATE["dirty"] = True
        return out
    if e["epoch"] != _PACK_STATE["epoch"] or e["version"] != reg._version:
        _refresh_all()
    return e["out"]


def pack_conv_train(weight, bias=None, stride=1, padding=0, dilation=1, cin_pad=None, cout_pad=None, transpose_flip=False,
                    cin_range=None):
    """pack_conv for the training step: the fp32 parameter is packed by one kernel.  transpose_flip=True gives the operand of
    the stride-1 dgrad conv (roles of cin / cout swapped, taps flipped, padding = dilation*(k-1) - padding given by the caller)."""
    w = weight.detach()
    assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()
    a, b, R, S = w.shape
    if cin_range is not None:
        b = cin_range[1]
    cout, cin = (b, a) if transpose_flip else (a, b)
    cout_pad = cout_pad or round_up(cout, 16)
    cin_pad = cin_pad or round_up(cin, 64)
    wp = _pack_device(w, cout_pad, cin_pad, 1 if transpose_flip else 0, cin_range)
    taps = [(r * dilation - padding, s * dilation - padding, r * S + s) for r in range(R) for s in range(S)]
    return PackedConv(wp, taps, 1, R * S, stride, 1, [0], [0], cout, _pad_bias(bias, cout_pad, w.device),
                      macs_per_pixel=R * S * cin * cout)


def pack_tapexp_train(weight, cin_pad, cp=4, transpose=False):
    """[co <= cp, ci, 3, 3] parameter -> 1x1 PackedConv of the tap-expanded conv (outputs t*cp + m, padded to 64) or, with
    transpose=True, of its dgrad (64 tap-expanded channels -> ci)."""
    w = weight.detach()
    assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and w.shape[2:] == (3, 3) and w.shape[0] <= cp
    co, ci = w.shape[:2]
    if transpose:
        wp = _pack_device(w, cin_pad, 64, 4 | (cp << 3))
        return PackedConv(wp, [(0, 0, 0)], 1, 1, 1, 1, [0], [0], ci, None, macs_per_pixel=9 * co * ci)
    wp = _pack_device(w, 64, cin_pad, 3 | (cp << 3))
    return PackedConv(wp, [(0, 0, 0)], 1, 1, 1, 1, [0], [0], 9 * cp, None, macs_per_pixel=9 * co * ci)


def pack_deconv8s4_train(weight, bias=None, cin_pad=None, cout_pad=None):
    """pack_deconv8s4 with the one-kernel packing (weight [Cin, Cout, 8, 8] of the transposed conv being evaluated)."""
    w = weight.detach()
    assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and w.shape[2:] == (8, 8)
    cin, cout = w.shape[:2]
    cout_pad = cout_pad or round_up(cout, 16)
    cin_pad = cin_pad or round_up(cin, 64)
    wp = _pack_device(w, cout_pad, cin_pad, 2)
    taps, ooh, oow = _DECONV_TAPS
    return PackedConv(wp, taps, 16, 4, 1, 4, ooh, oow, cout, _pad_bias(bias, cout_pad, w.device), macs_per_pixel=4 * cin * cout)


def _deconv_merged_tables():
    """Tap classes of the 8x8 / stride-4 / pad-2 transposed conv (see pack_deconv8s4): class (a, b) = output phases with
    rh in {2a, 2a+1}, rw in {2b, 2b+1}; sub-phase s = (rh % 2) * 2 + (rw % 2).  -> (taps [(dh, dw)] per class and tap,
    widx [(class * 4 + tap) * 4 + s], ooh / oow [class * 4 + s])."""
    def axis_taps(rho):
        return [(-1, rho + 6), (0, rho + 2)] if rho < 2 else [(0, rho + 2), (1, rho - 2)]
    taps, widx, ooh, oow = [], [], [], []
    for a in range(2):
        for b in range(2):
            for ti in range(2):
                for tj in range(2):
                    taps.append((axis_taps(2 * a)[ti][0], axis_taps(2 * b)[tj][0]))
                    for s in range(4):
                        rh, rw = 2 * a + s // 2, 2 * b + s % 2
                        assert axis_taps(rh)[ti][0] == taps[-1][0] and axis_taps(rw)[tj][0] == taps[-1][1]
                        widx.append(axis_taps(rh)[ti][1] * 8 + axis_taps(rw)[tj][1])
            for s in range(4):
                ooh.append(2 * a + s // 2)
                oow.append(2 * b + s % 2)
    return taps, widx, ooh, oow


_DECONV_MERGED = _deconv_merged_tables()
DECONV_MERGE = os.environ.get("CSBSR_DECONV_MERGE") == "1"


def _deconv_merge_ok(x, pc, y, rm, r0, r1, oh, ow):
    """The merged form needs the staged epilogue: bf16 NHWC output covering whole 128-pixel tiles, at most one residual."""
    if not (DECONV_MERGE and pc.nphases == 16 and pc.ntaps == 4 and pc.os == 4 and pc.cout_pad == 128 and isinstance(y, Fmap)):
        return False
    if pc.taps is not _DECONV_TAPS[0] and list(pc.taps) != list(_DECONV_TAPS[0]):
        return False
    if pc.cin_pad % 64 != 0 or rm is not None or (r0 is not None and r1 is not None):
        return False
    ih, iw = (oh, ow) if oh is not None else (x.h, x.w)
    th = 8 if iw > 8 else 16
    return y.h == 4 * ih and y.w == 4 * iw and ih % th == 0 and ih == x.h and iw == x.w


def _deconv_taps():
    def axis_taps(rho):
        return [(-1, rho + 6), (0, rho + 2)] if rho < 2 else [(0, rho + 2), (1, rho - 2)]
    taps, ooh, oow = [], [], []
    for rh in range(4):
        for rw in range(4):
            for dh, r in axis_taps(rh):
                for dw, s in axis_taps(rw):
                    taps.append((dh, dw, r * 8 + s))
            ooh.append(rh)
            oow.append(rw)
    return taps, ooh, oow


_DECONV_TAPS = _deconv_taps()


def pack_tapexp3x3(weight, cout_pad=None):
    """[co, ci, 3, 3] conv weight (padding 1) -> 1x1 PackedConv with 9*cp outputs (cp = co rounded up to 4 so one tap is
    a whole number of 8-byte words), channel t*cp + c = tap t of output c."""
    co, ci, R, S = weight.shape
    assert R == 3 and S == 3
    cp = round_up(co, 4)
    w = torch.zeros((9, cp, ci), dtype=torch.float32, device=weight.device)
    w[:, :co] = weight.detach().float().permute(2, 3, 0, 1).reshape(9, co, ci)
    pc = pack_conv(w.reshape(9 * cp, ci, 1, 1), cout_pad=cout_pad or round_up(9 * cp, 16))
    pc.macs_per_pixel = 9 * co * ci
    pc.tap_co = co
    return pc


def tap_gather3x3(z, co, out, r32=None):
    """out (PlanarWin or fp32 [N,co,H,W]) = r32 + sum over the 9 taps of the shifted tap-expanded map `z` (Fmap)."""
    ow = out if isinstance(out, PlanarWin) else PlanarWin(out, 0, co)
    rw = None if r32 is None else (r32 if isinstance(r32, PlanarWin) else PlanarWin(r32, 0, co))
    assert ow.c == co and (rw is None or rw.c == co) and z.c >= 9 * round_up(co, 4)
    _call("csbsr_tap_gather3x3", z.ptr(), z.pitch, z.coff, ow.t.data_ptr(), ow.t.shape[1], ow.coff,
          rw.t.data_ptr() if rw is not None else None, rw.t.shape[1] if rw is not None else 0, rw.coff if rw is not None else 0,
          z.n, z.h, z.w, co)
    return out


_WG_WS = {}


def wgrad(g, s, taps, stride=1, into=None):
    """Weight gradient wg[m][t][c] = sum_pix g[pix, m] * s[pix*stride + taps[t], c] (fp32 [round_up(g.c,128), T, s.c]).
    `g`, `s`: Fmaps whose channel windows are multiples of 64; `taps`: list of (dh, dw).  The pixel splits go through a
    workspace and a fixed-order reduction (no atomics).  into = (parameter, cin_range or None, cp): instead of returning wg the
    result is added to parameter.grad in the parameter's layout (cp > 0: tap-expanded accumulator) and None is returned."""
    d = _lib.WgradDesc()
    assert g.c % 64 == 0 and s.c % 64 == 0 and g.n == s.n and len(taps) <= _lib.MAX_TAPS
    d.g, d.n, d.gh, d.gw, d.g_pitch, d.g_coff, d.cg = g.ptr(), g.n, g.h, g.w, g.pitch, g.coff, g.c
    d.s, d.sh, d.sw, d.s_pitch, d.s_coff, d.cs = s.ptr(), s.h, s.w, s.pitch, s.coff, s.c
    d.ntaps, d.stride = len(taps), stride
    for i, (dh, dw) in enumerate(taps):
        d.dh[i], d.dw[i] = dh, dw
    L = _lib.lib()
    need = L.csbsr_conv_wgrad_workspace_bytes(C.byref(d))
    dev = g.t.device
    ws = _WG_WS.get(str(dev))
    if ws is None or ws.numel() < need:
        ws = _WG_WS[str(dev)] = torch.empty(max(need, 1 << 26), dtype=torch.uint8, device=dev)
    d.ws, d.ws_bytes = ws.data_ptr(), ws.numel()
    out = None
    if into is None:
        out = torch.empty((round_up(g.c, 128), len(taps), s.c), dtype=torch.float32, device=dev)
        d.wg = out.data_ptr()
    else:
        param, cin_range, cp = into
        a, btot = param.shape[:2]
        b0, b = cin_range if cin_range is not None else (0, btot)
        d.wg = None
        d.grad, d.grad_a, d.grad_b, d.grad_btot, d.grad_b0, d.grad_cp = param.grad.data_ptr(), a, b, btot, b0, cp
    if PROFILE_WG is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = L.csbsr_conv_wgrad(C.byref(d), _lib.stream_ptr())
    _lib.check(rc, "csbsr_conv_wgrad")
    _lib.count_launch("csbsr_conv_wgrad")
    if PROFILE_WG is not None:
        e1.record()
        PROFILE_WG.append(("wgrad n%d %dx%d cg%d cs%d taps%d s%d" % (g.n, g.h, g.w, g.c, s.c, len(taps), stride),
                           2.0 * g.n * g.h * g.w * g.c * s.c * len(taps), e0, e1))
    return out


def wgrad_target(weight):
    """The registered parameter whose .grad (a contiguous view of the optimizer's flat gradient) wgrad can accumulate into."""
    reg = registered(weight)
    return reg if (reg is not None and reg.grad is not None and reg.grad.is_contiguous()) else None


# ------------------------------------------------------------------ support-kernel launchers
def _call(name, *args):
    rc = getattr(_lib.lib(), name)(*args, _lib.stream_ptr())
    _lib.check(rc, name)
    _lib.count_launch(name)


def _f32(t):
    assert t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda
    return t.data_ptr()


def nchw_to_nhwc(x, y, cwrite=None):
    """fp32 [N,C,H,W] -> bf16 NHWC window `y` (channels [C, cwrite) zeroed)."""
    n, c, h, w = x.shape
    assert (y.n, y.h, y.w) == (n, h, w)
    _call("csbsr_nchw_f32_to_nhwc_bf16", _f32(x), y.ptr(), n, c, h, w, y.pitch, y.coff, cwrite or y.c)
    return y


def patchify(x, y, r, s, stride, pad, mean=None, rstd=None, clamp01=False):
    """fp32 [N,C,H,W] -> bf16 NHWC im2col rows ((r*S+s)*C+c ordering), see csbsr_patchify."""
    n, c, h, w = x.shape
    assert y.coff == 0 and y.n == n
    _call("csbsr_patchify", _f32(x), y.ptr(), n, c, h, w, y.h, y.w, r, s, stride, pad, y.pitch, y.c,
          _f32(mean) if mean is not None else None, _f32(rstd) if rstd is not None else None, int(clamp01))
    return y


def gap(x, out, c=None):
    c = c or x.c
    assert out.shape == (x.n, c)
    _call("csbsr_gap_nhwc", x.ptr(), _f32(out), x.n, x.h * x.w, x.pitch, x.coff, c)
    return out


def pack_chain(layers):
    """Weights of a fused conv chain (csrc/kpred_chain.cu) in its shared-memory plane layout: per layer
    [tap][cin_pad/8][cout_pad][8] bf16, concatenated.  `layers`: list of (weight [co,ci,R,S], cin_pad, cout_pad).
    Runs on the host (no device launches); returns a CPU uint8 tensor."""
    parts = []
    for w, cin_pad, cout_pad in layers:
        w = w.detach().float().cpu()
        co, ci, R, S = w.shape
        assert ci <= cin_pad and co <= cout_pad and cin_pad % 8 == 0
        a = torch.zeros((R * S, cin_pad, cout_pad), dtype=torch.float32)
        a[:, :ci, :co] = w.permute(2, 3, 1, 0).reshape(R * S, ci, co)
        a = a.view(R * S, cin_pad // 8, 8, cout_pad).permute(0, 1, 3, 2).contiguous()       # [tap][plane][n][8]
        parts.append(a.to(torch.bfloat16).view(torch.uint8).reshape(-1))
    return torch.cat(parts)


def kpred_sr_chain(img, wpack, out, slope=0.01):
    """fe_SR.0..4 of the kernel predictor in one kernel: img fp32 [B,3,H,W] -> out Fmap [B,H,W,64]."""
    n, c, h, w = img.shape
    assert c == 3 and out.pitch == 64 and out.coff == 0 and (out.n, out.h, out.w) == (n, h, w)
    assert wpack.numel() == _lib.lib().csbsr_kpred_wpack_bytes(0)
    rc = _lib.lib().csbsr_kpred_sr_chain(_f32(img), wpack.data_ptr(), out.ptr(), n, h, w, C.c_float(slope), _lib.stream_ptr())
    _lib.check(rc, "csbsr_kpred_sr_chain")
    _lib.count_launch("csbsr_kpred_sr_chain")
    return out


def kpred_cat_chain(x, wpack, cls_bias, out, ws, slope=0.01):
    """fe_cat.0..2 + global average pool in one kernel: x Fmap [B,H,W,64], cls_bias fp32 [B,5,5,64] -> out fp32 [B,c]."""
    assert x.pitch == 64 and x.coff == 0 and cls_bias.shape == (x.n, 5, 5, 64) and out.shape[0] == x.n
    assert wpack.numel() == _lib.lib().csbsr_kpred_wpack_bytes(1)
    rc = _lib.lib().csbsr_kpred_cat_chain(x.ptr(), wpack.data_ptr(), _f32(cls_bias), _f32(out), out.shape[1], ws.data_ptr(),
                                          ws.numel() * ws.element_size(), x.n, x.h, x.w, C.c_float(slope), _lib.stream_ptr())
    _lib.check(rc, "csbsr_kpred_cat_chain")
    _lib.count_launch("csbsr_kpred_cat_chain")
    return out


def kernel_update(v, pre, out, ke, ko, normalize=True):
    _call("csbsr_kernel_update", _f32(v), _f32(pre) if pre is not None else None, _f32(out), v.shape[0], ke, ko,
          int(normalize))
    return out


def vec_normalize(v, out):
    _call("csbsr_vec_normalize", _f32(v), _f32(out), v.shape[0], v.shape[1])
    return out


def broadcast_vec(v, y):
    _call("csbsr_broadcast_vec", _f32(v), y.ptr(), y.n, y.h * y.w, v.shape[1], y.pitch, y.coff, y.c)
    return y


def blur_per_sample(x, kvec, lr, err, ksize=21, stride=4):
    n, c, h, w = x.shape
    _call("csbsr_blur_per_sample", _f32(x), _f32(kvec), _f32(lr) if lr is not None else None, _f32(err), n, c, h, w,
          ksize, stride)
    return err


def bicubic_upsample(x, y, factor):
    n, c, h, w = x.shape
    assert y.shape == (n, c, h * factor, w * factor)
    _call("csbsr_bicubic_upsample", _f32(x), _f32(y), n * c, h, w, factor)
    return y


def clip_instnorm_stats(x, mean, rstd, do_clip=True, eps=1e-5):
    n, c, h, w = x.shape
    ws = torch.empty(_lib.lib().csbsr_instnorm_workspace_bytes(n * c) // 8, dtype=torch.float64, device=x.device)
    _call("csbsr_clip_instnorm_stats", _f32(x), _f32(mean), _f32(rstd), ws.data_ptr(), n * c, h * w, int(do_clip),
          C.c_float(eps))
    return mean, rstd


def maxpool3s2(x, y):
    _call("csbsr_maxpool3s2_nhwc", x.ptr(), y.ptr(), x.n, x.h, x.w, x.c, x.pitch, x.coff, y.pitch, y.coff)
    return y


def adaptive_avgpool(x, y, s):
    _call("csbsr_adaptive_avgpool_nhwc", x.ptr(), y.ptr(), x.n, x.h, x.w, s, x.c, x.pitch, x.coff, y.pitch, y.coff)
    return y


def bilinear(x, y, align_corners=False):
    assert x.c == y.c
    _call("csbsr_bilinear_nhwc", x.ptr(), y.ptr(), x.n, x.h, x.w, y.h, y.w, x.c, x.pitch, x.coff, y.pitch, y.coff,
          int(align_corners))
    return y


def bilinear_f32(x, y, align_corners=False):
    n, c, h, w = x.shape
    _call("csbsr_bilinear_f32", _f32(x), _f32(y), n * c, h, w, y.shape[2], y.shape[3], int(align_corners))
    return y


def bilinear_f32_sigmoid(x, y, align_corners=True):
    n, c, h, w = x.shape
    _call("csbsr_bilinear_f32_sigmoid", _f32(x), _f32(y), n * c, h, w, y.shape[2], y.shape[3], int(align_corners))
    return y


def bilinear_add(x, base, y, align_corners=True, relu=False):
    """y = [relu](base + bilinear(x)) (x is resampled to y's size; base / y may alias)."""
    assert x.c == y.c == base.c and (base.h, base.w) == (y.h, y.w)
    _call("csbsr_bilinear_add_nhwc", x.ptr(), base.ptr(), y.ptr(), x.n, x.h, x.w, y.h, y.w, x.c, x.pitch, x.coff,
          base.pitch, base.coff, y.pitch, y.coff, int(align_corners), int(relu))
    return y


def softmax_gather(logits, feats, ctx, c=None):
    """logits fp32 [N,1,H,W], feats Fmap -> ctx fp32 [N,C]."""
    c = c or feats.c
    _call("csbsr_softmax_gather", _f32(logits), feats.ptr(), _f32(ctx), feats.n, feats.h * feats.w, c, feats.pitch,
          feats.coff)
    return ctx
