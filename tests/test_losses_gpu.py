"""Parity of the loss kernels (C-ABI csbsr_sdf / csbsr_seg_loss / csbsr_seg_loss_wf_mean / csbsr_sr_loss) against the
golden vectors produced by the UNMODIFIED reference classes and against the oracle restatement.
Tolerances: SDF bit-exact (integer EDT, fp64 normalisation, fp32 store); loss values 2e-6 relative (fp64 accumulation
here vs fp32 reductions in torch); gradients 1e-6 abs + 1e-4 rel."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "losses.npz")


def test_sdf_bit_exact_vs_reference_golden_and_oracle():
    from csbsr_b200.engine import losses as LS
    from oracle import loss_ref as R
    g = np.load(GOLD)
    s = LS.compute_sdf(torch.from_numpy(g["mask"])).cpu().numpy()
    assert np.array_equal(s, g["sdf"])
    rng = np.random.default_rng(1)
    m = (rng.random((3, 1, 61, 45)) < 0.2).astype(np.float32)
    m[1] = 0                                              # empty mask -> zeros
    m[2, 0, 10:30, 5:20] = 1
    assert np.array_equal(LS.compute_sdf(torch.from_numpy(m)).cpu().numpy(), R.sdf(m).numpy())


def test_seg_loss_values_and_gradients_vs_reference_golden():
    from csbsr_b200.engine import losses as LS
    g = np.load(GOLD)
    pm, pa, m = (torch.from_numpy(g[k]) for k in ("p_main", "p_aux", "mask"))
    loss, gm, ga = LS.seg_loss(pm, pa, m, float(g["alpha"]), upstream=torch.from_numpy(g["plain_upstream"]), need_grad=True)
    assert np.allclose(loss.cpu().numpy(), g["plain_loss"], rtol=2e-6, atol=0)
    assert np.allclose(gm.cpu().numpy(), g["plain_grad_main"], rtol=1e-4, atol=1e-6)
    assert np.allclose(ga.cpu().numpy(), g["plain_grad_aux"], rtol=1e-4, atol=1e-6)
    assert (gm.cpu().numpy()[0, 0, :3, :5] == 0).all()        # below the clamp: no gradient


def test_wf_loss_mean_vs_reference_golden():
    from csbsr_b200.engine import losses as LS
    g = np.load(GOLD)
    pm, pa, m = (torch.from_numpy(g[k]) for k in ("p_main", "p_aux", "mask"))
    out = LS.seg_loss_wf_mean(pm, pa, m, float(g["alpha"]), 1.0).item()
    assert abs(out - float(g["map_mean"])) <= 2e-6 * abs(float(g["map_mean"]))


def test_kbpn_loss_vs_reference_golden():
    from csbsr_b200.engine import losses as LS
    g = np.load(GOLD)
    loss, kn = LS.kbpn_loss(*(torch.from_numpy(g[k]) for k in ("sr", "hr", "lr", "kvec", "kgt")))
    assert np.allclose(loss.cpu().numpy(), g["kbpn_loss"], rtol=2e-5, atol=1e-7)
    assert np.allclose(kn.cpu().numpy(), g["kbpn_kernel"], rtol=1e-5, atol=1e-9)
    total = LS.calc_loss(loss, torch.tensor(0.25, device=loss.device), 0.3).item()
    assert abs(total - (0.7 * float(g["kbpn_loss"].mean()) + 0.3 * 0.25)) < 1e-5


def test_wf_loss_fused_mean_and_gradient_vs_elementwise_autograd():
    """csbsr_seg_loss_wf_mean / _grad (closed form of the (B,B,H,W) mean) against the elementwise torch formulation."""
    from csbsr_b200.engine import losses as LS
    g = torch.Generator().manual_seed(21)
    B, H, W = 3, 40, 56
    mask = (torch.rand(B, 1, H, W, generator=g) > 0.8).float().cuda()
    pm0 = (0.02 + 0.96 * torch.rand(B, 1, H, W, generator=g)).cuda()
    pm0[0, 0, :2, :3] = 1e-9                                         # below the clamp: zero gradient there
    pa0 = (0.02 + 0.96 * torch.rand(B, 1, H, W, generator=g)).cuda()
    outs = []
    for fused in (True, False):
        pm, pa = pm0.clone().requires_grad_(True), pa0.clone().requires_grad_(True)
        lm = LS.seg_loss_train(pm, pa, mask, 0.37, wf_amp=1.0, fused=fused)
        assert tuple(lm.shape) == (B, B, H, W)
        (lm.mean() * 1.7).backward()
        outs.append((lm.mean().item(), pm.grad.clone(), pa.grad.clone()))
    (v1, gm1, ga1), (v0, gm0, ga0) = outs
    assert abs(v1 - v0) <= 2e-6 * abs(v0)
    assert (gm1 - gm0).abs().max().item() <= 1e-4 * gm0.abs().max().item()
    assert (ga1 - ga0).abs().max().item() <= 1e-4 * ga0.abs().max().item()
    assert gm1[0, 0, :2, :3].abs().max().item() == 0


@pytest.mark.parametrize("w_k", [0.0, 2.0])
def test_kbpn_loss_function_fwd_bwd_vs_torch_autograd(w_k):
    """engine/losses.py::_KBPNLossFn (csbsr_sr_loss forward, csbsr_sr_loss_bwd backward) against the elementwise torch form of
    KBPNLoss (sr_loss_functions.py:39-56) and its autograd, with a non-trivial upstream gradient per sample."""
    from csbsr_b200.engine.losses import _KBPNLossFn
    g = torch.Generator(device="cuda").manual_seed(31)
    B = 3
    sr = torch.rand(B, 3, 40, 56, device="cuda", generator=g)
    hr = torch.rand(B, 3, 40, 56, device="cuda", generator=g)
    hr[0, 0, :4] = sr[0, 0, :4]                               # exact ties: abs' subgradient 0 as in torch
    plr = torch.rand(B, 3, 10, 14, device="cuda", generator=g)
    lr = torch.rand(B, 3, 10, 14, device="cuda", generator=g)
    kn = torch.rand(B, 1, 21, 21, device="cuda", generator=g) / 441
    kg = torch.rand(B, 1, 21, 21, device="cuda", generator=g) / 441
    up = torch.tensor([0.3, -1.2, 2.0], device="cuda")
    a = [t.clone().requires_grad_(True) for t in (sr, plr, kn)]
    b = [t.clone().requires_grad_(True) for t in (sr, plr, kn)]
    loss = _KBPNLossFn.apply(a[0], hr, a[1], lr, a[2], kg, 0.4, 0.4, w_k)
    ref = 0.4 * (b[0] - hr).abs().mean((1, 2, 3)) + 0.4 * (b[1] - lr).abs().mean((1, 2, 3)) + w_k * ((b[2] - kg) ** 2).mean((1, 2, 3))
    (loss * up).sum().backward()
    (ref * up).sum().backward()
    assert torch.allclose(loss, ref, rtol=2e-6, atol=1e-7)
    assert torch.allclose(a[0].grad, b[0].grad, rtol=1e-6, atol=1e-10) and torch.allclose(a[1].grad, b[1].grad, rtol=1e-6, atol=1e-10)
    if w_k != 0:
        assert torch.allclose(a[2].grad, b[2].grad, rtol=1e-5, atol=1e-10)
    else:
        assert a[2].grad is None
