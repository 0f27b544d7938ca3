"""Time the joint training step (config #3: batch 8 per GPU, 224^2 HR crops, iteration 40000, w^F on) on one GPU."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--layers", action="store_true")
    ap.add_argument("--detector", default="PSPNet", choices=["PSPNet", "HRNet_OCR"])
    ap.add_argument("--graph", action="store_true", help="capture forward + loss + backward in a CUDA graph")
    a = ap.parse_args()
    from csbsr_b200 import _lib
    from csbsr_b200.config import cfg
    from csbsr_b200.engine.losses import calc_loss
    from csbsr_b200.engine.optim import FusedAdam
    from csbsr_b200.modeling import params as P
    from csbsr_b200.modeling.build_model import JointModelWithLoss
    c = cfg.clone()
    c.merge_from_file("config/config_csbsr_pspnet.yaml")
    c.SOLVER.SEG_FAIL_ORIENTED_WEIGHT4SS_AMP = 1.0
    if a.detector == "HRNet_OCR":                              # config #4: HRNet-W48 + OCR, beta = 0.9
        c.MODEL.DETECTOR_TYPE = "HRNet_OCR"
        c.SOLVER.TASK_LOSS_WEIGHT = 0.9
    m = JointModelWithLoss(c, num_train_ds=1000, resume_iter=40000)
    sd = P.synth_state_dict(P.kbpn_param_shapes(), prefix="sr_model.")
    seg_shapes = P.hrnet_ocr_param_shapes() if a.detector == "HRNet_OCR" else P.pspnet_param_shapes()
    sd.update(P.synth_state_dict(seg_shapes, prefix="segmentation_model."))
    m.load_state_dict(sd)
    m.cuda().train()
    opt = FusedAdam(m.parameters(), lr=c.SOLVER.LR)
    g = torch.Generator().manual_seed(1)
    B, S = a.batch, a.size
    hr = torch.rand(B, 3, S, S, generator=g).cuda()
    lr = torch.nn.functional.interpolate(hr, size=(S // 4, S // 4), mode="bicubic", antialias=True).clamp(0, 1)
    mask = (torch.rand(B, 1, S, S, generator=g) > 0.9).float().cuda()
    kgt = torch.rand(B, 1, 21, 21, generator=g).cuda()
    kgt = kgt / kgt.sum(dim=(2, 3), keepdim=True)

    def step(it):
        seg_loss, sr_loss, *_ = m(it, lr, sr_targets=hr, segment_targets=mask, kernel_targets=kgt)
        loss = calc_loss(sr_loss, seg_loss.mean(), c.SOLVER.TASK_LOSS_WEIGHT)
        loss.backward()
        opt.step()
        return loss

    if a.graph:
        m.dropout = True
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(3):
                l = step(40000 + i)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        static_loss = None
        with torch.cuda.graph(graph):
            seg_loss, sr_loss, *_ = m(40100, lr, sr_targets=hr, segment_targets=mask, kernel_targets=kgt)
            static_loss = calc_loss(sr_loss, seg_loss.mean(), c.SOLVER.TASK_LOSS_WEIGHT)
            static_loss.backward()

        def step(it):                                     # noqa: F811
            graph.replay()
            opt.step()
            return static_loss
    for i in range(a.warmup):
        l = step(40000 + i)
    torch.cuda.synchronize()
    print("warm loss", l.item(), "mem GB", torch.cuda.max_memory_allocated() / 2 ** 30)
    if a.profile:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step(40010)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=90))
    if a.layers:
        import collections
        from csbsr_b200 import kernels as K
        K.PROFILE, K.PROFILE_WG = [], []
        step(40015)
        torch.cuda.synchronize()
        for name, rec in (("conv fwd+dgrad", K.PROFILE), ("wgrad", K.PROFILE_WG)):
            agg = collections.OrderedDict()
            for r in rec:
                a_ = agg.setdefault(r[0], [0, 0.0, 0.0])
                a_[0] += 1; a_[1] += r[2].elapsed_time(r[3]); a_[2] += r[1]
            tot = sum(v[1] for v in agg.values())
            print("== %s: %.2f ms, %.0f TFLOP/s (padded)" % (name, tot, sum(v[2] for v in agg.values()) / tot / 1e9))
            for k, (n, t, f) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
                print("   %-52s x%3d %7.3f ms %5.1f%% %7.1f TF/s" % (k, n, t, 100 * t / tot, f / t / 1e9))
        K.PROFILE = K.PROFILE_WG = None
    n0 = _lib.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        l = step(40020 + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print("train step %.1f ms  -> %.2f steps/s, %.1f img/s (batch %d, %dx%d)  loss %.4f  launches/step %s" %
          (ms, 1000 / ms, B * 1000 / ms, B, S, S, l.item(), (_lib.LAUNCHES - n0) // a.steps))


if __name__ == "__main__":
    main()
