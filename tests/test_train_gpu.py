"""Training-path kernels: conv weight / data gradients on the tcgen05 engine against
torch autograd in fp32 on the same bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rand_nhwc(n, h, w, c, cpad, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.zeros(n, h, w, cpad, dtype=torch.bfloat16)
    t[..., :c] = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16)
    return t.cuda()


WGRAD_CASES = [
    # (n, h, w, cin, cout, k, stride, pad, dil)
    (2, 20, 24, 64, 64, 3, 1, 1, 1),
    (2, 17, 9, 128, 192, 3, 1, 2, 2),
    (1, 33, 40, 64, 128, 3, 2, 1, 1),
    (2, 16, 16, 256, 64, 1, 1, 0, 1),
    (1, 32, 48, 128, 128, 8, 4, 2, 1),
    (3, 8, 8, 320, 64, 3, 1, 1, 1),
]


@pytest.mark.parametrize("case", WGRAD_CASES)
def test_conv_wgrad_vs_autograd(case):
    from csbsr_b200 import kernels as K
    n, h, w, ci, co, k, st, pad, dil = case
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    x = _rand_nhwc(n, h, w, ci, K.round_up(ci, 64), 1)
    oh = (h + 2 * pad - dil * (k - 1) - 1) // st + 1
    ow = (w + 2 * pad - dil * (k - 1) - 1) // st + 1
    dy = _rand_nhwc(n, oh, ow, co, K.round_up(co, 64), 2)
    taps = [(r * dil - pad, s * dil - pad) for r in range(k) for s in range(k)]
    wg = K.wgrad(K.Fmap(dy), K.Fmap(x), taps, stride=st)
    got = wg[:co, :, :ci].reshape(co, k, k, ci).permute(0, 3, 1, 2)
    wt = torch.zeros(co, ci, k, k, device="cuda", requires_grad=True)
    y = F.conv2d(x[..., :ci].permute(0, 3, 1, 2).float(), wt, stride=st, padding=pad, dilation=dil)
    y.backward(dy[..., :co].permute(0, 3, 1, 2).float())
    ref = wt.grad
    err = (got - ref).abs().max().item()
    print("wgrad", case, "max err", err, "ref max", ref.abs().max().item())
    assert err <= 2e-3 * ref.abs().max().item() + 1e-3


def test_deconv_wgrad_vs_autograd():
    """ConvTranspose2d(8, 4, 2): G = layer input, S = dL/dy."""
    from csbsr_b200 import kernels as K
    torch.backends.cudnn.allow_tf32 = False
    n, h, w, ci, co = 2, 12, 16, 128, 64
    x = _rand_nhwc(n, h, w, ci, 128, 3)
    dy = _rand_nhwc(n, 4 * h, 4 * w, co, 64, 4)
    taps = [(r - 2, s - 2) for r in range(8) for s in range(8)]
    wg = K.wgrad(K.Fmap(x), K.Fmap(dy), taps, stride=4)
    got = wg[:ci, :, :co].reshape(ci, 8, 8, co).permute(0, 3, 1, 2)
    wt = torch.zeros(ci, co, 8, 8, device="cuda", requires_grad=True)
    y = F.conv_transpose2d(x.permute(0, 3, 1, 2).float(), wt, stride=4, padding=2)
    y.backward(dy.permute(0, 3, 1, 2).float())
    err = (got - wt.grad).abs().max().item()
    print("deconv wgrad max err", err, "ref max", wt.grad.abs().max().item())
    assert err <= 2e-3 * wt.grad.abs().max().item() + 1e-3


CONV_CASES = [
    # (n, h, w, cin, cout, k, stride, pad, dil, bias)
    (2, 20, 24, 64, 49, 3, 1, 1, 1, True),
    (2, 24, 16, 128, 128, 3, 1, 2, 2, False),
    (1, 32, 40, 64, 128, 3, 2, 1, 1, False),
    (2, 16, 24, 64, 128, 1, 2, 0, 1, False),
    (1, 32, 48, 128, 128, 8, 4, 2, 1, False),
    (2, 30, 22, 3, 64, 7, 2, 3, 1, False),
    (2, 12, 12, 569, 569, 3, 1, 1, 1, True),
]


def _rel(a, b):
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_autograd_fn(case):
    """y, dx, dw, db of the tcgen05 conv Function against torch autograd in fp32 (bf16-rounded operands)."""
    from csbsr_b200 import autograd as A
    n, h, w, ci, co, k, st, pad, dil, has_bias = case
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(5)
    x0 = torch.randn(n, ci, h, w, generator=g).to(torch.bfloat16).float().cuda()
    wt = (torch.randn(co, ci, k, k, generator=g) * (2.0 / (ci * k * k)) ** 0.5).to(torch.bfloat16).float().cuda()
    b = torch.randn(co, generator=g).cuda() if has_bias else None
    x_ref = x0.clone().requires_grad_(True)
    w_ref = wt.clone().requires_grad_(True)
    b_ref = b.clone().requires_grad_(True) if has_bias else None
    y_ref = F.conv2d(x_ref, w_ref, b_ref, stride=st, padding=pad, dilation=dil)
    up = torch.randn(y_ref.shape, generator=torch.Generator().manual_seed(6)).to(torch.bfloat16).float().cuda()
    y_ref.backward(up)

    x = x0.clone().requires_grad_(True)
    w_ = wt.clone().requires_grad_(True)
    b_ = b.clone().requires_grad_(True) if has_bias else None
    y = A.conv2d(A.to_nhwc(x), w_, b_, stride=st, padding=pad, dilation=dil)
    assert y.shape[3] == A.cpad(co) and (y[..., co:] == 0).all()
    y.backward(A.to_nhwc(up))
    print("conv fn", case, "y", _rel(A.to_nchw(y, co), y_ref), "dx", _rel(x.grad, x_ref.grad), "dw", _rel(w_.grad, w_ref.grad))
    assert _rel(A.to_nchw(y, co).detach(), y_ref.detach()) <= 1e-2
    assert _rel(x.grad, x_ref.grad) <= 1e-2
    assert _rel(w_.grad, w_ref.grad) <= 1e-2
    if has_bias:
        assert _rel(b_.grad, b_ref.grad) <= 1e-2


def test_deconv8s4_autograd_fn():
    from csbsr_b200 import autograd as A
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(7)
    for ci, co in ((128, 128), (3, 128)):
        x0 = torch.randn(2, ci, 10, 12, generator=g).to(torch.bfloat16).float().cuda()
        wt = (torch.randn(ci, co, 8, 8, generator=g) * 0.05).to(torch.bfloat16).float().cuda()
        x_ref, w_ref = x0.clone().requires_grad_(True), wt.clone().requires_grad_(True)
        y_ref = F.conv_transpose2d(x_ref, w_ref, stride=4, padding=2)
        up = torch.randn(y_ref.shape, generator=g).to(torch.bfloat16).float().cuda()
        y_ref.backward(up)
        x, w_ = x0.clone().requires_grad_(True), wt.clone().requires_grad_(True)
        y = A.deconv8s4(A.to_nhwc(x), w_)
        y.backward(A.to_nhwc(up))
        print("deconv fn", ci, co, _rel(A.to_nchw(y, co), y_ref), _rel(x.grad, x_ref.grad), _rel(w_.grad, w_ref.grad))
        assert _rel(A.to_nchw(y, co).detach(), y_ref.detach()) <= 1e-2
        assert _rel(x.grad, x_ref.grad) <= 1e-2 and _rel(w_.grad, w_ref.grad) <= 1e-2
