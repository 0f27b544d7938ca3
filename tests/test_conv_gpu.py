"""Parity of the tcgen05 implicit-GEMM conv (C-ABI csbsr_conv_igemm) against torch fp32 convs on the
same bf16-rounded operands.  Tolerance: bf16 output rounding (rel 2^-8) + fp32 accumulation order."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16).float()


def _check(got, ref, tol=2e-2):
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= tol * scale, "max err %g vs scale %g" % (err, scale)


@pytest.mark.parametrize("n,cin,cout,h,w,k,stride,pad,dil", [
    (2, 64, 128, 16, 16, 1, 1, 0, 1),
    (1, 128, 64, 24, 40, 3, 1, 1, 1),
    (2, 192, 256, 16, 32, 3, 1, 1, 1),
    (1, 64, 48, 56, 56, 3, 1, 2, 2),
    (1, 64, 64, 28, 28, 3, 1, 4, 4),
    (2, 128, 128, 64, 64, 8, 4, 2, 1),
    (1, 64, 128, 32, 32, 3, 2, 1, 1),
    (1, 64, 64, 32, 48, 1, 2, 0, 1),
    (3, 64, 16, 5, 5, 3, 1, 1, 1),
    (1, 576, 576, 16, 16, 3, 1, 1, 1),
])
def test_conv_matches_torch(n, cin, cout, h, w, k, stride, pad, dil):
    from csbsr_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(1)
    x = _bf(torch.randn(n, cin, h, w, device="cuda", generator=g))
    wt = _bf(torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5)
    b = torch.randn(cout, device="cuda", generator=g)
    ref = F.conv2d(x, wt, b, stride=stride, padding=pad, dilation=dil)
    pc = K.pack_conv(wt, b, stride=stride, padding=pad, dilation=dil)
    xf = K.Fmap.from_nchw(x)
    y = K.Fmap.empty(n, ref.shape[2], ref.shape[3], K.round_up(cout, 16))
    K.conv(xf, pc, y)
    torch.cuda.synchronize()
    _check(y.to_nchw_f32(cout), ref)


@pytest.mark.parametrize("n,cin,cout,h,w,k,pad,kpad,pitch", [
    (2, 32, 32, 24, 40, 3, 1, 32, 64),      # 64-byte swizzle rows, input stored with a 64-channel pitch
    (1, 32, 49, 56, 56, 3, 1, 32, 32),
    (2, 27, 49, 16, 24, 1, 0, 32, 64),
    (2, 3, 64, 20, 20, 3, 1, 16, 16),       # 32-byte swizzle rows
    (1, 16, 128, 9, 33, 3, 1, 16, 64),
])
def test_conv_narrow_k_chunks(n, cin, cout, h, w, k, pad, kpad, pitch):
    """K chunks of 32 / 16 channels (64B / 32B swizzle) for the layers whose cin is 32 or less."""
    from csbsr_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(4)
    x = _bf(torch.randn(n, cin, h, w, device="cuda", generator=g))
    wt = _bf(torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5)
    ref = F.conv2d(x, wt, None, padding=pad)
    pc = K.pack_conv(wt, None, padding=pad, cin_pad=kpad)
    assert pc.cin_pad == kpad
    xf = K.Fmap.from_nchw(x, cpad=pitch)
    xf.t[..., cin:] = 7.0                     # channels past cin_pad must never be read; inside it they meet zero weights
    y = K.Fmap.empty(n, h, w, K.round_up(cout, 16))
    K.conv(xf.window(0, kpad), pc, y)
    torch.cuda.synchronize()
    _check(y.to_nchw_f32(cout), ref)


def test_deconv8s4_matches_torch():
    from csbsr_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(2)
    n, cin, cout, h, w = 2, 128, 128, 16, 32
    x = _bf(torch.randn(n, cin, h, w, device="cuda", generator=g))
    wt = _bf(torch.randn(cin, cout, 8, 8, device="cuda", generator=g) / (cin * 4) ** 0.5)
    ref = F.conv_transpose2d(x, wt, None, stride=4, padding=2)
    pc = K.pack_deconv8s4(wt)
    y = K.Fmap.empty(n, 4 * h, 4 * w, cout)
    K.conv(K.Fmap.from_nchw(x), pc, y)
    torch.cuda.synchronize()
    _check(y.to_nchw_f32(cout), ref)


@pytest.mark.parametrize("n,cin,h,w,res,act", [(2, 128, 16, 32, None, False), (3, 128, 8, 24, "r1", True), (1, 64, 16, 8, "r1m", True),
                                               (2, 192, 24, 40, "r0", True), (2, 128, 16, 16, "bias", True)])
def test_deconv8s4_merged_subphases(n, cin, h, w, res, act, monkeypatch):
    """The merged form of the 8x8 / stride-4 transposed conv (4 tap classes x 4 sub-phases, N = 256 tiles, both epilogue
    teams on every tile; csbsr_conv_desc.nsub) against torch and against the 16-phase form (same products; the taps are
    accumulated in a different order, so single bf16 roundings may differ)."""
    from csbsr_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(12)
    cout = 128
    x = _bf(torch.randn(n, cin, h, w, device="cuda", generator=g))
    wt = _bf(torch.randn(cin, cout, 8, 8, device="cuda", generator=g) / (cin * 4) ** 0.5)
    bias = torch.randn(cout, device="cuda", generator=g) if res == "bias" else None
    r = _bf(torch.randn(n, cout, 4 * h, 4 * w, device="cuda", generator=g)) if res in ("r0", "r1", "r1m") else None
    ref = F.conv_transpose2d(x, wt, bias, stride=4, padding=2)
    if res == "r0":
        ref = ref + r
    if act:
        ref = F.leaky_relu(ref, 0.1)
    if res == "r1":
        ref = ref + r
    if res == "r1m":
        ref = ref - r
    pc = K.pack_deconv8s4(wt, bias)
    kw = dict(act=K.ACT_LEAKY if act else K.ACT_NONE, slope=0.1)
    if res == "r0":
        kw["r0"] = K.Fmap.from_nchw(r)
    if res in ("r1", "r1m"):
        kw["r1"] = K.Fmap.from_nchw(r)
        kw["r1_sign"] = 1.0 if res == "r1" else -1.0
    xf = K.Fmap.from_nchw(x)
    monkeypatch.setattr(K, "DECONV_MERGE", True)        # opt-in (CSBSR_DECONV_MERGE=1): measured no faster than 16 phases
    assert K._deconv_merge_ok(xf, pc, K.Fmap.empty(n, 4 * h, 4 * w, cout), None, kw.get("r0"), kw.get("r1"), None, None)
    y = K.conv(xf, pc, K.Fmap.empty(n, 4 * h, 4 * w, cout), **kw)
    torch.cuda.synchronize()
    _check(y.to_nchw_f32(cout), ref)
    monkeypatch.setattr(K, "DECONV_MERGE", False)
    y16 = K.conv(xf, pc, K.Fmap.empty(n, 4 * h, 4 * w, cout), **kw)
    torch.cuda.synchronize()
    a, b = y.t.float(), y16.t.float()
    assert (a - b).abs().max().item() <= 2 ** -7 * ref.abs().max().item() and (a != b).float().mean().item() < 0.05


def test_epilogue_variants():
    from csbsr_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(3)
    n, cin, cout, h, w = 2, 64, 128, 16, 16
    x = _bf(torch.randn(n, cin, h, w, device="cuda", generator=g))
    wt = _bf(torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (cin * 9) ** 0.5)
    b = torch.randn(cout, device="cuda", generator=g)
    r0 = _bf(torch.randn(n, cout, h, w, device="cuda", generator=g))
    rm = _bf(torch.randn(n, cout, h, w, device="cuda", generator=g))
    r1 = _bf(torch.randn(n, cout, h, w, device="cuda", generator=g))
    base = F.conv2d(x, wt, b, padding=1)
    pc = K.pack_conv(wt, b, padding=1)
    xf = K.Fmap.from_nchw(x)
    # pre-activation residual + relu (ResNet BasicBlock, extractors.py:52-70)
    y = K.Fmap.empty(n, h, w, cout)
    K.conv(xf, pc, y, act=K.ACT_RELU, r0=K.Fmap.from_nchw(r0))
    _check(y.to_nchw_f32(), F.relu(base + r0))
    # leaky + post subtract (UpBlock l0 - x, kbpn.py:467-468)
    K.conv(xf, pc, y, act=K.ACT_LEAKY, slope=0.25, r1=K.Fmap.from_nchw(r1), r1_sign=-1.0)
    _check(y.to_nchw_f32(), F.leaky_relu(base, 0.25) - r1)
    # SFT combine: features * sigmoid(conv) + shift (kbpn.py:515-518)
    K.conv(xf, pc, y, act=K.ACT_SIGMOID, rm=K.Fmap.from_nchw(rm), r1=K.Fmap.from_nchw(r1))
    _check(y.to_nchw_f32(), torch.sigmoid(base) * rm + r1)
    # write into a channel slice of a wider buffer (torch.cat replacement)
    wide = K.Fmap.empty(n, h, w, 3 * cout, zero=True)
    K.conv(xf, pc, wide.window(cout, cout))
    torch.cuda.synchronize()
    _check(wide.t[..., cout:2 * cout].permute(0, 3, 1, 2).float(), base)
    assert wide.t[..., :cout].abs().max().item() == 0 and wide.t[..., 2 * cout:].abs().max().item() == 0


def test_f32_planar_output_and_class_bias():
    from csbsr_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(4)
    n, cin, h, w = 2, 128, 24, 24
    x = _bf(torch.randn(n, cin, h, w, device="cuda", generator=g))
    wt = _bf(torch.randn(3, cin, 3, 3, device="cuda", generator=g) / (cin * 9) ** 0.5)
    r32 = torch.randn(n, 3, h, w, device="cuda", generator=g)
    ref = F.conv2d(x, wt, None, padding=1) + r32
    pc = K.pack_conv(wt, None, padding=1)
    y = torch.empty(n, 3, h, w, device="cuda")
    K.conv(K.Fmap.from_nchw(x), pc, y, r32=r32)
    torch.cuda.synchronize()
    _check(y, ref, tol=1e-3)
    # per-sample border-class bias (5x5 classes): spatially constant conditioning folded into the bias
    cout = 32
    wt2 = _bf(torch.randn(cout, cin, 1, 1, device="cuda", generator=g) / cin ** 0.5)
    cb = torch.randn(n, 25, cout, device="cuda", generator=g).contiguous()
    pc2 = K.pack_conv(wt2)
    y2 = K.Fmap.empty(n, h, w, cout)
    K.conv(K.Fmap.from_nchw(x), pc2, y2, bias=cb, bias_sn=25 * cout, bias_sc=cout, cls_bw=2)
    torch.cuda.synchronize()
    def cls(i, m):
        return i if i < 2 else (2 if i < m - 2 else 4 - (m - 1 - i))
    ci = torch.tensor([cls(i, h) for i in range(h)], device="cuda")
    cj = torch.tensor([cls(j, w) for j in range(w)], device="cuda")
    idx = ci[:, None] * 5 + cj[None, :]
    ref2 = F.conv2d(x, wt2) + cb[:, idx].permute(0, 3, 1, 2)
    _check(y2.to_nchw_f32(cout), ref2)


@pytest.mark.parametrize("co", [3, 6, 12])
def test_tap_expansion_3x3_few_channels(co):
    """128 -> co 3x3 conv as a 1x1 GEMM with N = 9*co plus the shifted gather, accumulated into fp32 planes."""
    from csbsr_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(9)
    n, ci, h, w = 2, 128, 24, 40
    x = _bf(torch.randn(n, ci, h, w, device="cuda", generator=g))
    wt = _bf(torch.randn(co, ci, 3, 3, device="cuda", generator=g) / (ci * 9) ** 0.5)
    acc = torch.randn(n, co + 2, h, w, device="cuda", generator=g)
    ref = F.conv2d(x, wt, None, padding=1) + acc[:, 1:1 + co]
    pc = K.pack_tapexp3x3(wt)
    z = K.conv(K.Fmap.from_nchw(x), pc, K.Fmap.empty(n, h, w, pc.cout_pad))
    out = torch.zeros(n, co + 2, h, w, device="cuda")
    K.tap_gather3x3(z, co, K.PlanarWin(out, 1, co), K.PlanarWin(acc, 1, co))
    torch.cuda.synchronize()
    assert (out[:, 0] == 0).all() and (out[:, co + 1] == 0).all()
    _check(out[:, 1:1 + co], ref)


def test_conv_cluster_multicast_mode(monkeypatch):
    """CSBSR_CLUSTER=2: CTA pairs on neighbouring pixel tiles, each loading half of every weight tile with TMA multicast
    (odd tile count -> one dummy tile; per-sample class bias; staged and direct epilogues)."""
    from csbsr_b200 import kernels as K
    monkeypatch.setenv("CSBSR_CLUSTER", "2")
    g = torch.Generator(device="cuda").manual_seed(12)
    for (n, cin, cout, h, w) in ((3, 128, 128, 24, 40), (1, 192, 208, 40, 24)):
        x = _bf(torch.randn(n, cin, h, w, device="cuda", generator=g))
        wt = _bf(torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (cin * 9) ** 0.5)
        b = torch.randn(cout, device="cuda", generator=g)
        ref = F.conv2d(x, wt, b, padding=1)
        y = K.Fmap.empty(n, h, w, K.round_up(cout, 16))
        K.conv(K.Fmap.from_nchw(x), K.pack_conv(wt, b, padding=1), y)
        torch.cuda.synchronize()
        _check(y.to_nchw_f32(cout), ref)
