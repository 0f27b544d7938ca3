"""Fused kernel-predictor chains (csrc/kpred_chain.cu) against a plain PyTorch fp32 restatement of the same layers
(reference model/modeling/kbpn.py:528-541, 562-578) on the same bf16-rounded operands: every layer output is rounded to
bf16 exactly where the kernel rounds it, so only the fp32 accumulation order differs (tolerance: 1 bf16 ulp of the
largest activation for the maps, 1e-3 relative for the pooled vector)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.to(torch.bfloat16).float()


def _weights(gen):
    rn = lambda *s: torch.randn(*s, generator=gen)
    return {"sr0": rn(49, 3, 3, 3) * 0.25, "sr1": rn(32, 49, 1, 1) * 0.2, "sr2": rn(32, 32, 3, 3) * 0.08,
            "sr3": rn(32, 32, 3, 3) * 0.08, "sr4": rn(49, 32, 3, 3) * 0.08, "cat0": rn(32, 49, 1, 1) * 0.2,
            "cat1": rn(32, 32, 3, 3) * 0.08, "cat2": rn(49, 32, 3, 3) * 0.08}


def _sr_ref(x, w):
    a = _bf(F.relu(F.conv2d(_bf(x), _bf(w["sr0"]), padding=1)))
    a = _bf(F.leaky_relu(F.conv2d(a, _bf(w["sr1"])), 0.01))
    a = _bf(F.leaky_relu(F.conv2d(a, _bf(w["sr2"]), padding=1), 0.01))
    a = _bf(F.leaky_relu(F.conv2d(a, _bf(w["sr3"]), padding=1), 0.01))
    return _bf(F.leaky_relu(F.conv2d(a, _bf(w["sr4"]), padding=1), 0.01))


def _cls_map(cb, H, W):
    """[B,5,5,64] border-class table -> [B,64,H,W] (class 0,1 | 2 interior | 3,4 at the far border)."""
    def cls(n):
        i = torch.arange(n)
        c = torch.full((n,), 2, dtype=torch.long)
        c[i < 2] = i[i < 2]
        far = (n - 1 - i) < 2
        c[far] = 4 - (n - 1 - i[far])
        return c
    cy, cx = cls(H).to(cb.device), cls(W).to(cb.device)
    return cb[:, cy][:, :, cx].permute(0, 3, 1, 2)


def _cat_ref(a, cb, w):
    B, _, H, W = a.shape
    z = F.conv2d(a, _bf(w["cat0"])) + _cls_map(cb, H, W)[:, :32]
    z = _bf(F.leaky_relu(z, 0.01))
    z = _bf(F.leaky_relu(F.conv2d(z, _bf(w["cat1"]), padding=1), 0.01))
    z = F.conv2d(z, _bf(w["cat2"]), padding=1)
    return z.mean(dim=(2, 3))


@pytest.mark.parametrize("B,H,W", [(2, 40, 56), (1, 130, 250), (3, 96, 128), (1, 33, 123), (2, 448, 448)])
def test_kpred_chains_vs_torch(B, H, W):
    from csbsr_b200 import kernels as K
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    gen = torch.Generator().manual_seed(100 + H)
    w = {k: v.cuda() for k, v in _weights(gen).items()}
    x = torch.rand(B, 3, H, W, generator=gen).cuda()
    as1x1 = lambda t: t.permute(0, 2, 3, 1).reshape(t.shape[0], -1, 1, 1)
    wsr = K.pack_chain([(as1x1(w["sr0"]), 32, 64), (w["sr1"], 64, 32), (w["sr2"], 32, 32), (w["sr3"], 32, 32),
                        (w["sr4"], 32, 64)]).cuda()
    wcat = K.pack_chain([(w["cat0"], 64, 32), (w["cat1"], 32, 32), (w["cat2"], 32, 64)]).cuda()
    out = K.Fmap(torch.full((B, H, W, 64), 7.0, dtype=torch.bfloat16, device="cuda"))
    K.kpred_sr_chain(x, wsr, out, slope=0.01)
    torch.cuda.synchronize()
    ref = _sr_ref(x, w)
    got = out.to_nchw_f32(49)
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    print("sr chain", (B, H, W), "max err", err, "max |ref|", scale)
    assert err <= scale * 2.0 ** -7                                    # one bf16 ulp at the largest magnitude
    assert (out.t[..., 49:].float() == 0).all()                         # padded output channels are written as zeros
    # cat chain on the chain's own output (so both kernels see identical inputs)
    cb = (torch.randn(B, 5, 5, 64, generator=gen) * 0.1).cuda()
    gap = torch.zeros(B, 49, device="cuda")
    nws = K._lib.lib().csbsr_kpred_workspace_bytes(B, H, W)
    ws = torch.zeros(nws // 4, device="cuda")
    K.kpred_cat_chain(out, wcat, cb, gap, ws, slope=0.01)
    torch.cuda.synchronize()
    gref = _cat_ref(got, cb, w)
    gerr = (gap - gref).abs().max().item()
    print("cat chain gap max err", gerr, "max |ref|", gref.abs().max().item())
    assert gerr <= 1e-3 * max(gref.abs().max().item(), 1e-3) + 1e-5
    # determinism: the pooled sums are reduced in a fixed order
    gap2 = torch.zeros_like(gap)
    K.kpred_cat_chain(out, wcat, cb, gap2, ws, slope=0.01)
    assert torch.equal(gap, gap2)


def test_kbpn_fused_predictor_matches_layerwise_path():
    """KBPNEngine with the fused chains vs the one-launch-per-layer path (both CUDA): same bf16 rounding points except that
    the fused chain pools fe_cat.2's fp32 accumulators (the layer-wise path pools its bf16-rounded map), so the blur-kernel
    refinement differs by bf16 rounding noise (kernel vector <= 0.5 %, SR image <= 1e-2 on a [0,1] image)."""
    from tests.test_model_gpu import _model_and_sd
    m, sd = _model_and_sd()
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 3, 24, 40, generator=g).cuda()
    sr_eng, _ = m._ensure_engines(torch.device("cuda", 0))
    sr_eng.fused_kpred = True
    sr_a, kv_a = sr_eng.forward(x)
    sr_a, kv_a = sr_a.clone(), kv_a.clone()
    sr_eng.fused_kpred = False
    sr_b, kv_b = sr_eng.forward(x)
    torch.cuda.synchronize()
    print("fused vs layerwise: sr", (sr_a - sr_b).abs().max().item(), "kvec", (kv_a - kv_b).abs().max().item())
    assert (sr_a - sr_b).abs().max().item() <= 1e-2
    assert (kv_a - kv_b).abs().max().item() <= 5e-3 * kv_b.abs().max().item()
