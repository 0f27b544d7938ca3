"""Glue ops of the training step (csbsr_b200/glue.py on csrc/glue.cu) against the aten ops the reference's train step uses at
the same places (F.interpolate / adaptive_avg_pool2d / max_pool2d / leaky_relu / cat / instance_norm / Dropout2d ...), forward
and backward, on the same bf16-rounded inputs.  Tolerance: bf16 output rounding (2^-8 relative) unless stated."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).cuda()


def _nchw(x):
    return x.permute(0, 3, 1, 2).float()


def _close(a, b, tol=1e-2, what=""):
    err = (a.float() - b.float()).abs().max().item()
    ref = b.float().abs().max().item() + 1e-12
    assert err <= tol * ref, (what, err, ref)


def _run(fn_ours, fn_ref, inputs, seed=99, tol=1e-2):
    """inputs: NHWC bf16 leaves.  Compares outputs and the gradient of every input under a random upstream gradient."""
    xs = [x.clone().requires_grad_(True) for x in inputs]
    rs = [x.clone().float().requires_grad_(True) for x in inputs]
    y = fn_ours(*xs)
    r = fn_ref(*rs)
    _close(y, r, tol, "forward")
    up = _rnd(tuple(y.shape), seed)
    y.backward(up)
    r.backward(up.float())
    for i, (a, b) in enumerate(zip(xs, rs)):
        _close(a.grad, b.grad, tol, "grad of input %d" % i)


def test_add_sub_relu_sft():
    from csbsr_b200 import glue as G
    a, b, c = _rnd((2, 9, 11, 64), 1), _rnd((2, 9, 11, 64), 2), _rnd((2, 9, 11, 64), 3)
    _run(G.add, lambda x, y: x + y, [a, b])
    _run(G.sub, lambda x, y: x - y, [a, b])
    _run(lambda x, y: G.add(x, y, relu=True), lambda x, y: F.relu(x + y), [a, b])
    _run(G.relu, F.relu, [a])
    _run(lambda x: G.leaky_relu(x, 0.1), lambda x: F.leaky_relu(x, 0.1), [a])
    _run(G.sft_combine, lambda f, s, t: f * torch.sigmoid(s) + t, [a, b, c])


def test_concat_with_real_channels():
    from csbsr_b200 import glue as G
    parts = [_rnd((2, 6, 7, 64), 4), _rnd((2, 6, 7, 128), 5), _rnd((2, 6, 7, 192), 6)]
    real = (48, 96, 192)

    def ref(*ps):
        cat = torch.cat([p[..., :r] for p, r in zip(ps, real)], dim=3)
        return F.pad(cat, (0, 384 - cat.shape[3]))
    _run(lambda *ps: G.concat(ps, real), ref, parts)
    _run(lambda *ps: G.concat(ps), lambda *ps: torch.cat(ps, dim=3), parts)


@pytest.mark.parametrize("h,w,oh,ow,align", [(7, 9, 14, 18, False), (1, 1, 28, 20, False), (3, 3, 28, 28, False), (6, 6, 28, 28, False),
                                             (7, 5, 28, 20, True), (14, 14, 56, 56, True), (5, 6, 5, 6, False)])
def test_bilinear_fwd_bwd(h, w, oh, ow, align):
    from csbsr_b200 import glue as G
    x = _rnd((2, h, w, 64), 7)
    ref = lambda t: F.interpolate(t.permute(0, 3, 1, 2), size=(oh, ow), mode="bilinear", align_corners=align).permute(0, 2, 3, 1)
    _run(lambda t: G.bilinear(t, (oh, ow), align), ref, [x])


@pytest.mark.parametrize("h,w,s", [(28, 28, 1), (28, 28, 2), (28, 28, 3), (28, 28, 6), (16, 20, 3), (7, 9, 6)])
def test_adaptive_avgpool_fwd_bwd(h, w, s):
    from csbsr_b200 import glue as G
    x = _rnd((2, h, w, 128), 8)
    ref = lambda t: F.adaptive_avg_pool2d(t.permute(0, 3, 1, 2), (s, s)).permute(0, 2, 3, 1)
    _run(lambda t: G.adaptive_avgpool(t, s), ref, [x])


@pytest.mark.parametrize("h,w", [(16, 24), (15, 17), (56, 56)])
def test_maxpool_fwd_bwd_with_ties(h, w):
    from csbsr_b200 import glue as G
    x = F.relu(_rnd((2, h, w, 64), 9).float()).to(torch.bfloat16)          # post-ReLU maps: whole windows tie at 0
    ref = lambda t: F.max_pool2d(t.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    _run(G.maxpool3s2, ref, [x])


def test_gap_and_expand_classes():
    from csbsr_b200 import glue as G
    x = _rnd((3, 10, 12, 64), 10)
    xs, rs = x.clone().requires_grad_(True), x.clone().float().requires_grad_(True)
    y, r = G.gap(xs, 49), rs[..., :49].mean(dim=(1, 2))
    _close(y, r, 1e-2, "gap")
    up = torch.randn(3, 49, device="cuda")
    y.backward(up)
    r.backward(up)
    _close(xs.grad, rs.grad, 1e-2, "gap grad")
    for bw, (h, w) in ((1, (9, 13)), (2, (12, 40)), (2, (5, 6))):
        k = 2 * bw + 1
        small = _rnd((2, k, k, 64), 11)
        idx = lambda n: torch.tensor([i if i < bw else (2 * bw - (n - 1 - i) if i >= n - bw else bw) for i in range(n)], device="cuda")
        _run(lambda t: G.expand_classes(t, h, w, bw), lambda t: t[:, idx(h)][:, :, idx(w)], [small], tol=2e-2)


def test_layout_conversions_and_instance_norm():
    from csbsr_b200 import glue as G
    g = torch.Generator().manual_seed(12)
    img = torch.rand(2, 3, 20, 24, generator=g).cuda()
    a, r = img.clone().requires_grad_(True), img.clone().requires_grad_(True)
    y = G.to_nhwc(a)
    assert y.shape == (2, 20, 24, 64) and (y[..., 3:] == 0).all()
    _close(y[..., :3], r.permute(0, 2, 3, 1), 1e-2, "to_nhwc")
    back = G.to_nchw(y, 3)
    _close(back, img, 1e-2, "to_nchw")
    up = torch.randn(2, 3, 20, 24, device="cuda")
    back.backward(up)
    _close(a.grad, up, 1e-2, "round-trip gradient")
    a, r = img.clone().requires_grad_(True), img.clone().requires_grad_(True)
    y, yr = G.instance_norm(a), F.instance_norm(r, eps=1e-5)
    assert (y - yr).abs().max().item() <= 2e-5
    y.backward(up)
    yr.backward(up)
    assert (a.grad - r.grad).abs().max().item() <= 2e-4 * r.grad.abs().max().item() + 1e-6


def test_dropout2d_statistics_and_backward():
    from csbsr_b200 import glue as G
    st = G.DropoutState(torch.device("cuda", 0), seed=7)
    x = torch.ones(16, 4, 4, 256, dtype=torch.bfloat16, device="cuda").requires_grad_(True)
    st.begin_step()
    y = G.dropout2d(x, 200, 0.3, st)
    vals = y[:, 0, 0, :].float()
    assert (vals[:, 200:] == 0).all()                                     # padding channels stay zero
    kept = vals[:, :200] != 0
    assert abs(kept.float().mean().item() - 0.7) < 0.05
    assert torch.allclose(vals[:, :200][kept], torch.tensor(1 / 0.7, device="cuda"), rtol=1e-2)
    assert (y == y[:, :1, :1, :]).all()                                    # whole channels are dropped
    y.backward(torch.ones_like(y))
    assert torch.equal(x.grad, y.detach())
    st.begin_step()
    y2 = G.dropout2d(x.detach(), 200, 0.3, st)
    assert not torch.equal(y2, y.detach())                                # a new step draws a new mask
    st2 = G.DropoutState(torch.device("cuda", 0), seed=7)
    st2.begin_step()
    assert torch.equal(G.dropout2d(x.detach(), 200, 0.3, st2), y.detach())   # same seed + step + layer -> same mask


def test_bias_grad_and_conv_act_epilogue():
    from csbsr_b200 import autograd as A, glue as G
    dy = _rnd((3, 17, 19, 128), 13)
    got = G.bias_grad(dy, 100)
    _close(got, dy[..., :100].float().sum(dim=(0, 1, 2)), 1e-3, "bias grad")
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(14)
    x0 = torch.randn(2, 64, 12, 16, generator=g).to(torch.bfloat16).float().cuda()
    w0 = (torch.randn(96, 64, 3, 3, generator=g) * 0.05).to(torch.bfloat16).float().cuda()
    b0 = torch.randn(96, generator=g).cuda()
    for act, ref_act in (("relu", F.relu), (("lrelu", 0.1), lambda t: F.leaky_relu(t, 0.1))):
        x, w_, b_ = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
        xr, wr, br = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
        y = A.conv2d(A.to_nhwc(x), w_, b_, padding=1, act=act)
        yr = ref_act(F.conv2d(xr, wr, br, padding=1))
        _close(A.to_nchw(y, 96), yr, 1e-2, "conv+act")
        up = torch.randn(yr.shape, generator=torch.Generator().manual_seed(15)).to(torch.bfloat16).float().cuda()
        y.backward(A.to_nhwc(up))
        yr.backward(up)
        _close(x.grad, xr.grad, 1e-2, "dx")
        _close(w_.grad, wr.grad, 1e-2, "dw")
        _close(b_.grad, br.grad, 1e-2, "db")


@pytest.mark.parametrize("ci,h,w", [(128, 12, 16), (512, 9, 11)])
def test_tap_expanded_conv3x3_fwd_bwd(ci, h, w):
    """sr_reconst / output_conv as tap-expanded 1x1 GEMMs against F.conv2d (3 outputs, padding 1), forward, dx and dw."""
    from csbsr_b200 import autograd as A
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(21)
    x0 = torch.randn(2, ci, h, w, generator=g).to(torch.bfloat16).float().cuda()
    w0 = (torch.randn(3, ci, 3, 3, generator=g) * (2.0 / (9 * ci)) ** 0.5).to(torch.bfloat16).float().cuda()
    x, w_ = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
    xr, wr = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
    y = A.conv3x3_few_outputs(A.to_nhwc(x), w_)
    yr = F.conv2d(xr, wr, None, padding=1)
    assert (y[..., 3:] == 0).all()
    _close(A.to_nchw(y, 3), yr, 2e-2, "tapexp forward")          # z is rounded to bf16 before the nine taps are summed
    up = torch.randn(yr.shape, generator=torch.Generator().manual_seed(22)).to(torch.bfloat16).float().cuda()
    y.backward(A.to_nhwc(up))
    yr.backward(up)
    _close(x.grad, xr.grad, 1e-2, "tapexp dx")
    _close(w_.grad, wr.grad, 1e-2, "tapexp dw")


def test_patch_split_join_and_crop_flip_vs_oracle(tmp_path):
    """Input pipeline kernels (SplitPatch / JointPatch, crop + flips) bit-equal to the oracle restatement of the reference's
    host code (oracle/patch_ref.py, pinned to model/data/samplers/patch_sampler.py on CPU)."""
    from csbsr_b200.data import augment as AU
    from csbsr_b200.data.samplers.patch_sampler import JointPatch, SplitPatch
    from csbsr_b200.utils import save_output as SO
    from oracle import patch_ref as PR
    g = torch.Generator().manual_seed(31)
    x = torch.rand(3, 224, 336, generator=g)
    patches, shape = SplitPatch(7, 3, 112, 112)(x)
    p_ref, s_ref = PR.split_patch(x, 7, 3, 112, 112)
    assert torch.equal(patches.cpu(), p_ref) and list(shape) == list(s_ref)
    seg = torch.rand(2 * 6, 1, 448, 448, generator=g)                  # two images' x4 segmentation outputs
    sh = shape.copy(); sh[[5, 6]] = sh[[5, 6]] * 4; sh[[1, 4]] = 1
    assert torch.equal(JointPatch()(seg, sh).cpu(), PR.joint_patch(seg, sh))
    rng = np.random.default_rng(5)
    imgs = [(rng.random((300 + 7 * i, 280 + 5 * i, 3)) * 255).astype(np.uint8) for i in range(4)]
    masks = [(rng.random(im.shape[:2]) > 0.5).astype(np.uint8) * 255 for im in imgs]
    prm = AU.draw_params([im.shape[:2] for im in imgs], (224, 224), np.random.default_rng(9))
    out = AU.crop_flip_batch(imgs, prm, (224, 224)).cpu()
    outm = AU.crop_flip_batch(masks, prm, (224, 224)).cpu()
    for i, (im, mk) in enumerate(zip(imgs, masks)):
        y0, x0, hf, vf = [int(v) for v in prm[i]]
        assert torch.equal(out[i], PR.crop_flip(im, y0, x0, hf, vf, 224, 224))
        assert torch.equal(outm[i], PR.crop_flip(mk, y0, x0, hf, vf, 224, 224))
    assert prm[:, 2].max() == 1 or prm[:, 3].max() == 1                # some flip drawn with this seed

    class A:
        output_dirname = str(tmp_path)
    SO.save_img(str(tmp_path), out[:2], ["a.png", "b.png"])
    SO.save_mask(A, outm[:2], ["a.png", "b.png"], 0.5)
    SO.save_kernel(A, torch.rand(2, 1, 21, 21), ["a.png", "b.png"], 2)
    from PIL import Image
    back = np.asarray(Image.open(tmp_path / "images" / "a.png"))
    assert back.shape == (224, 224, 3) and np.array_equal(back, (out[0] * 255).byte().permute(1, 2, 0).numpy())
    assert (tmp_path / "masks" / "th_0.50" / "b.png").exists() and (tmp_path / "kernels_origin" / "b_0_origin.png").exists()


def test_device_prefetcher_order_and_contents():
    """data/prefetch.py: batches arrive in order with the right contents while the next copy is already in flight."""
    from csbsr_b200.data.prefetch import DevicePrefetcher
    host = [(torch.full((3, 64, 64), float(i)).pin_memory(), torch.arange(16, dtype=torch.int32).add(i).pin_memory()) for i in range(5)]
    got = []
    for a, b in DevicePrefetcher(iter(host), torch.device("cuda", 0)):
        x = a * 2.0                                    # some work on the current stream that reads the buffers
        got.append((x.sum().item(), int(b[0].item())))
    assert got == [(2.0 * i * 3 * 64 * 64, i) for i in range(5)]
