"""Achieved bandwidth of the HBM-class kernels against their ALGORITHMIC bytes,
CUDA-event timed on the launching stream, batch-64 inputs (larger than the 126 MB L2 wherever the kernel's working set is).

  python scripts/bench_hbm_kernels.py            -> table + one JSON line (committed under profiles/)
  bench.py imports run() for its `roofline_hbm` block.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6554.9


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def noisy_crack_case(b, size=448, seed=0):
    """SURVEY.md App. E perf fixture: crack |y - (0.5 x + 60 + 10 sin(x/17))| < 3 with a noisy probability map (every
    threshold has border corners: ~190 EDT evaluations per image over the 99 thresholds)."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(size, dtype=torch.float32), torch.arange(size, dtype=torch.float32), indexing="ij")
    masks, probs = [], []
    for i in range(b):
        m = ((yy - (0.5 * xx + 60 + 10 * i + 10 * torch.sin(xx / 17))).abs() < 3).float()
        soft = torch.nn.functional.avg_pool2d(m[None, None], 5, 1, 2)[0, 0]
        p = (soft * 0.9 + 0.08 * torch.randn(size, size, generator=g)).clamp(0, 1)
        masks.append(m)
        probs.append(p)
    return torch.stack(probs)[:, None].contiguous(), torch.stack(masks)[:, None].contiguous()


def run(B=64, quiet=True):
    from csbsr_b200 import kernels as K, _lib
    from csbsr_b200.data import degrade as G
    from csbsr_b200.engine import inference as E, losses as LS
    from csbsr_b200.utils import synth
    peak = _peak()
    rows = []

    def report(name, ms, nbytes, units, note=""):
        gbs = nbytes / ms / 1e6
        rows.append({"kernel": name, "ms": ms, "algorithmic_bytes": int(nbytes), "gbs": gbs, "frac": gbs / peak, "units": units,
                     "note": note})
        if not quiet:
            print("%-44s %8.3f ms  %8.1f GB/s  %5.1f %% of %.0f  %s" % (name, ms, gbs, 100 * gbs / peak, peak, note))

    hr, mask = synth.batch(0, 8, 448)
    hr = hr.repeat(B // 8, 1, 1, 1).cuda()
    mask = mask.repeat(B // 8, 1, 1, 1).cuda()
    params = torch.as_tensor(synth.degradation_params(B)).cuda()
    ms = timed(lambda: G.degrade(hr, params))
    report("csbsr_degrade_fused", ms, B * 2560740, "%d x 448^2 images" % B,
           "composed 36x36/s4 kernels: 97.5 MFLOP/img -> %.1f TFLOP/s fp32 (FP32-FMA-bound, not HBM-bound)" % (B * 97.5e6 / (ms * 1e-3) / 1e12))
    ms3 = timed(lambda: G.degrade(hr, params, return_blurred=True))
    report("csbsr_degrade (3 launches, blurred in HBM)", ms3, B * 2560740, "%d x 448^2 images" % B,
           "direct 21x21 blur: 531 MFLOP/img -> %.1f TFLOP/s fp32" % (B * 0.531e9 / (ms3 * 1e-3) / 1e12))
    sr = torch.rand(B, 3, 448, 448, device="cuda")
    mean = torch.empty(B * 3, device="cuda")
    rstd = torch.empty(B * 3, device="cuda")
    ms = timed(lambda: K.clip_instnorm_stats(sr, mean, rstd, do_clip=True))
    report("csbsr_clip_instnorm_stats", ms, B * 2 * 2408448, "%d x 3 x 448^2 fp32" % B, "read + clamped write + stats")
    x = K.Fmap.empty(8, 448, 448, 128)
    x.t.normal_()
    y = K.Fmap.empty(8, 448, 448, 128)
    slope = torch.tensor([0.2], device="cuda")
    ms = timed(lambda: _lib.lib().csbsr_prelu_fwd(x.ptr(), y.ptr(), slope.data_ptr(), x.t.numel(), _lib.stream_ptr()))
    report("csbsr_prelu_fwd", ms, 2 * x.t.numel() * 2, "8 x 448^2 x 128 bf16")
    gv = torch.empty(8, 128, device="cuda")
    ms = timed(lambda: K.gap(x, gv, 128))
    report("csbsr_gap_nhwc", ms, x.t.numel() * 2, "8 x 448^2 x 128 bf16")
    n = 89_100_000 // 4 * 4
    p, g_, m, v = (torch.zeros(n, device="cuda") for _ in range(4))
    ms = timed(lambda: _lib.lib().csbsr_adam_step(p.data_ptr(), g_.data_ptr(), m.data_ptr(), v.data_ptr(), n, 2e-5, 0.9, 0.999,
                                                  1e-8, 3, 1.0, 1, _lib.stream_ptr()))
    report("csbsr_adam_step", ms, n * 28, "89.1 M parameters", "p,g,m,v read; p,m,v written")
    del p, g_, m, v, x, y
    prob = torch.rand(B, 1, 448, 448, device="cuda")
    ms = timed(lambda: [E.seg_metrics(prob[i:i + 16], mask[i:i + 16], with_hd=False, to_host=False) for i in range(0, B, 16)], reps=3)
    report("csbsr_seg_metrics (AIU only)", ms, B * 1605632, "%d images x 99 thresholds" % B)
    pn, mn = noisy_crack_case(16)
    pn, mn = pn.cuda(), mn.cuda()
    ms = timed(lambda: E.seg_metrics(pn, mn, with_hd=True, to_host=False), reps=3)
    report("csbsr_seg_metrics (AIU + HD/MSD sweep), App. E noisy crack", ms, 16 * 1605632, "16 images x 99 thresholds",
           "%.3f ms / image; %.0f (image x threshold) EDT + percentile evaluations per s -- integer / latency-bound on-chip work, "
           "the HBM fraction is the algorithmic 1.6 MB / image only" % (ms / 16, 16 * 99 / ms * 1e3))
    ms = timed(lambda: LS.compute_sdf(mask[:16]), reps=3)
    report("csbsr_sdf", ms, 16 * 2 * 802816, "16 masks 448^2")
    return {"peak_gbs": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy bandwidth)", "kernels": rows}


if __name__ == "__main__":
    out = run(quiet=False)
    print(json.dumps(out))
