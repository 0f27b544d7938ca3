"""End-to-end parity of the sm_100a JointModel against the fp32 oracle restatement (oracle/torch_ref.py,
itself pinned to the unmodified reference by tests/golden) on the same synthetic weights and inputs.

Tolerances (bf16 activations/weights, fp32 accumulation, ~100 conv layers deep):
  SR image: max-abs <= 3e-2 on a [0,1] image and PSNR(ours, oracle) >= 40 dB;
  segmentation probability: max-abs <= 2e-2, mean-abs <= 5e-3 (measured 0.012 / 0.002);  blur kernel: max-abs <= 2% of max|kernel| (the kernel is renormalised by its own
  sum at every stage, kbpn.py:391-392, which amplifies rounding when that sum is far from 1)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _model_and_sd(blur_skip=False):
    from csbsr_b200.config import cfg
    from csbsr_b200.modeling.build_model import JointModel
    from csbsr_b200.modeling import params as P
    c = cfg.clone()
    c.merge_from_file("config/config_csbsr_pspnet.yaml")
    if blur_skip:
        c.MODEL.DETECTOR_TYPE = "PSPNet_BlurSkip"
    m = JointModel(c)
    sd = P.synth_state_dict(P.kbpn_param_shapes(), prefix="sr_model.")
    sd.update(P.synth_state_dict(P.pspnet_param_shapes(blur_dim=441 if blur_skip else None), prefix="segmentation_model."))
    m.load_state_dict(sd, strict=True)
    return m, sd


def _psnr(a, b):
    mse = ((a - b) ** 2).mean().item()
    return 10 * math.log10(1.0 / max(mse, 1e-20))


def test_kbpn_stagewise_vs_oracle():
    from oracle import torch_ref as T
    m, sd = _model_and_sd()
    sdc = {k: v.cuda() for k, v in sd.items()}
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 3, 24, 32, generator=g).cuda()
    sr_eng, _ = m._ensure_engines(torch.device("cuda", 0))
    sr, kvec = sr_eng.forward(x)
    with torch.no_grad():
        sr_ref, kvec_ref = T.kbpn_forward(sdc, x)
    torch.cuda.synchronize()
    print("sr max-abs", (sr - sr_ref).abs().max().item(), "psnr", _psnr(sr, sr_ref),
          "kvec max-abs", (kvec - kvec_ref.view(2, -1)).abs().max().item())
    assert (sr - sr_ref).abs().max().item() <= 3e-2
    assert _psnr(sr, sr_ref) >= 40.0
    assert (kvec - kvec_ref.view(2, -1)).abs().max().item() <= 2e-2 * kvec_ref.abs().max().item()


def test_pspnet_vs_oracle():
    from oracle import torch_ref as T
    m, sd = _model_and_sd()
    sdc = {k: v.cuda() for k, v in sd.items()}
    g = torch.Generator().manual_seed(6)
    img = torch.randn(2, 3, 96, 128, generator=g).cuda()
    _, ss_eng = m._ensure_engines(torch.device("cuda", 0))
    seg, aux = ss_eng.forward(img)
    with torch.no_grad():
        seg_ref, aux_ref = T.pspnet_forward(sdc, img)
    torch.cuda.synchronize()
    print("seg max-abs", (seg - seg_ref).abs().max().item(), "aux max-abs", (aux - aux_ref).abs().max().item())
    assert (seg - seg_ref).abs().max().item() <= 2e-2
    assert (aux - aux_ref).abs().max().item() <= 4e-2          # auxiliary head (training loss only, not an eval output): measured 0.028
    assert (seg - seg_ref).abs().mean().item() <= 5e-3


@pytest.mark.parametrize("b,h,w", [(3, 24, 32), (1, 40, 24)])
def test_joint_model_vs_oracle(b, h, w):
    from oracle import torch_ref as T
    m, sd = _model_and_sd()
    sdc = {k: v.cuda() for k, v in sd.items()}
    g = torch.Generator().manual_seed(7)
    x = torch.rand(b, 3, h, w, generator=g)
    m.chunk = 2
    sr, seg, kp, aux = m(x, torch.zeros(b, 1, 7, 7), return_aux=True)
    with torch.no_grad():
        sr_ref, seg_ref, kp_ref, aux_ref = T.joint_forward(sdc, x.cuda())
    torch.cuda.synchronize()
    assert sr.shape == (b, 3, 4 * h, 4 * w) and seg.shape == (b, 1, 4 * h, 4 * w) and kp.shape == (b, 1, 21, 21)
    print("sr", (sr - sr_ref).abs().max().item(), _psnr(sr, sr_ref), "seg", (seg - seg_ref).abs().max().item(),
          (seg - seg_ref).abs().mean().item(), "kp", (kp - kp_ref).abs().max().item())
    assert (sr - sr_ref).abs().max().item() <= 3e-2 and _psnr(sr, sr_ref) >= 40.0
    assert (seg - seg_ref).abs().max().item() <= 2e-2 and (seg - seg_ref).abs().mean().item() <= 5e-3
    assert (kp - kp_ref).abs().max().item() <= 2e-2 * kp_ref.abs().max().item()
    assert sr.min().item() >= 0.0 and sr.max().item() <= 1.0


def test_blurskip_joint_model_vs_oracle_and_golden():
    """PSPNet_BlurSkip (config #5 detector): p + BlurSkip(p, kernel) with the 441 conditioning channels folded."""
    import os
    import numpy as np
    from oracle import torch_ref as T
    m, sd = _model_and_sd(blur_skip=True)
    sdc = {k: v.cuda() for k, v in sd.items()}
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "joint_blurskip.npz"))
    x = torch.from_numpy(g["x"])
    sr, seg, kp = m(x, torch.zeros(x.shape[0], 1, 7, 7))
    with torch.no_grad():
        sr_ref, seg_ref, kp_ref, _ = T.joint_forward(sdc, x.cuda(), blur_skip=True)
    torch.cuda.synchronize()
    print("blurskip sr", (sr - sr_ref).abs().max().item(), "seg", (seg - seg_ref).abs().max().item(), (seg - seg_ref).abs().mean().item())
    assert (sr - sr_ref).abs().max().item() <= 3e-2 and _psnr(sr, sr_ref) >= 40.0
    assert (seg - seg_ref).abs().max().item() <= 2e-2 and (seg - seg_ref).abs().mean().item() <= 5e-3
    # and against the unmodified reference's own outputs (fp16-stored fixture)
    assert np.abs(seg.cpu().numpy() - g["seg"].astype(np.float32)).max() <= 2e-2
    assert np.abs(sr.cpu().numpy() - g["sr"].astype(np.float32)).max() <= 3e-2


def test_baseline_shape_joint_model_eager_graph_and_metrics():
    """The benchmarked path itself (BASELINE config #2): 40 synthetic 448^2 crack images degraded on the device,
    JointModel at 112^2 -> 448^2 with KBPN chunks of 16 (+ a ragged 8) and segmentation chunk 32 (+ a ragged 8), as bench.py runs it;
    the same batch through chunks of 8 / 24 must give bit-identical images, maps and kernels.
      * eager outputs vs the fp32 oracle (TF32 off) on the first and last four images, tolerances of this file;
      * the CUDA-graph replay bench.py times must reproduce the eager outputs bit for bit;
      * AIU counts / HD / MSD of the device sweep on the model's own probability maps equal the oracle's sweep on the same maps."""
    import numpy as np
    from oracle import metrics_ref as M
    from oracle import torch_ref as T
    from csbsr_b200.data import degrade as G
    from csbsr_b200.engine import inference as E
    from csbsr_b200.utils import synth
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    m, sd = _model_and_sd()
    sdc = {k: v.cuda() for k, v in sd.items()}
    B = 40
    hr_u, mask_u = synth.batch(0, 8, 448)
    hr = hr_u.repeat(5, 1, 1, 1).cuda()
    mask = mask_u.repeat(5, 1, 1, 1).cuda()
    params = torch.as_tensor(synth.degradation_params(B, seed=5)).cuda()
    m.chunk, m.seg_chunk = 16, 32                      # bench.py's configuration: KBPN chunks of 16, 16 and a ragged 8
    lr, _ = G.degrade(hr, params)
    sr, seg, kp = m(lr, None)
    sr, seg, kp = sr.clone(), seg.clone(), kp.clone()
    torch.cuda.synchronize()
    # an image's result must not depend on which other images share its launches (tile scheduling, cta_group::2 rule)
    m.chunk, m.seg_chunk = 8, 24
    sr8, seg8, kp8 = m(lr, None)
    assert torch.equal(sr8, sr) and torch.equal(seg8, seg) and torch.equal(kp8, kp), "results depend on the chunking"
    m.chunk, m.seg_chunk = 16, 32
    for sl in (slice(0, 4), slice(B - 4, B)):
        with torch.no_grad():
            sr_ref, seg_ref, kp_ref, _ = T.joint_forward(sdc, lr[sl])
        e_sr = (sr[sl] - sr_ref).abs().max().item()
        e_seg, m_seg = (seg[sl] - seg_ref).abs().max().item(), (seg[sl] - seg_ref).abs().mean().item()
        print("448^2", sl, "sr", e_sr, _psnr(sr[sl], sr_ref), "seg", e_seg, m_seg, "kp", (kp[sl] - kp_ref).abs().max().item())
        assert e_sr <= 3e-2 and _psnr(sr[sl], sr_ref) >= 40.0
        assert e_seg <= 2e-2 and m_seg <= 5e-3
        assert (kp[sl] - kp_ref).abs().max().item() <= 2e-2 * kp_ref.abs().max().item()
        del sr_ref, seg_ref
    # graph replay of the same forward (what bench.py times)
    static_lr = lr.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        m(static_lr, None)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        g_sr, g_seg, g_kp = m(static_lr, None)
    g_sr.zero_(); g_seg.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(g_sr, sr) and torch.equal(g_seg, seg) and torch.equal(g_kp, kp)
    # metric sweep on the network's own maps: bit-exact vs the oracle sweep (two images: full oracle HD is ~4 s each)
    r = E.seg_metrics(seg[:16], mask[:16], with_hd=True, percent=50)
    seg_h, mask_h = seg[:2].cpu().numpy(), mask[:2].cpu().numpy()
    inter, union = M.iou_counts(seg[:16].cpu().numpy(), mask[:16].cpu().numpy())
    hd, msd = M.distance_metrics(seg_h, mask_h, 50)
    assert np.array_equal(r["inter"], inter) and np.array_equal(r["union"], union)
    assert np.array_equal(r["hd"][:2], hd) and np.array_equal(r["msd"][:2], msd)


def _hrnet_model_and_sd():
    from csbsr_b200.config import cfg
    from csbsr_b200.modeling.build_model import JointModel
    from csbsr_b200.modeling import params as P
    c = cfg.clone()
    c.merge_from_file("config/config_csbsr_pspnet.yaml")
    c.MODEL.DETECTOR_TYPE = "HRNet_OCR"
    c.SOLVER.TASK_LOSS_WEIGHT = 0.9
    m = JointModel(c)
    sd = P.synth_state_dict(P.kbpn_param_shapes(), prefix="sr_model.")
    sd.update(P.synth_state_dict(P.hrnet_ocr_param_shapes(), prefix="segmentation_model."))
    m.load_state_dict(sd, strict=True)
    return m, sd


def test_hrnet_ocr_vs_oracle():
    """HRNet-W48 + OCR head (config #4 detector) on a normalised image, against the fp32 oracle."""
    from oracle import torch_ref as T
    m, sd = _hrnet_model_and_sd()
    sdc = {k: v.cuda() for k, v in sd.items()}
    g = torch.Generator().manual_seed(8)
    img = torch.randn(2, 3, 128, 160, generator=g).cuda()
    _, ss_eng = m._ensure_engines(torch.device("cuda", 0))
    seg, aux = ss_eng.forward(img)
    with torch.no_grad():
        seg_ref, aux_ref = T.hrnet_ocr_forward(sdc, img)
    torch.cuda.synchronize()
    print("hrnet seg", (seg - seg_ref).abs().max().item(), (seg - seg_ref).abs().mean().item(), "aux", (aux - aux_ref).abs().max().item())
    assert (seg - seg_ref).abs().mean().item() <= 5e-3 and (aux - aux_ref).abs().mean().item() <= 5e-3
    assert (seg - seg_ref).abs().max().item() <= 2e-2 and (aux - aux_ref).abs().max().item() <= 2e-2


def test_hrnet_joint_model_vs_reference_golden():
    """JointModel(KBPN + HRNet_OCR) against the unmodified reference's outputs (tests/golden/joint_hrnet.npz)."""
    import os
    import numpy as np
    m, sd = _hrnet_model_and_sd()
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "joint_hrnet.npz"))
    x = torch.from_numpy(g["x"])
    sr, seg, kp = m(x, torch.zeros(x.shape[0], 1, 7, 7))
    assert sr.shape == (2, 3, 64, 96) and seg.shape == (2, 1, 64, 96) and kp.shape == (2, 1, 21, 21)
    d = np.abs(seg.cpu().numpy() - g["seg"].astype(np.float32))
    print("hrnet joint seg", d.max(), d.mean())
    assert d.max() <= 2e-2 and d.mean() <= 5e-3
    assert np.abs(sr.cpu().numpy() - g["sr"].astype(np.float32)).max() <= 3e-2
