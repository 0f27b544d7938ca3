#!/usr/bin/env python
"""Benchmark of the CSBSR hot path (BASELINE.json metric): SR + segmentation + AIU/AHD95 images/sec.

One "step" = one pass of the whole eval hot path over one batch of synthetic 448x448 crack images per GPU:
on-the-fly degradation -> KBPN x4 blind SR -> clip + instance norm -> PSPNet -> AIU + HD/MSD sweeps (99
thresholds).  Workload = BASELINE.json configs[1] ("CSBSR w/ PSPNet x4 inference bf16 batch 64 on 1 B200 with
on-the-fly anisotropic-blur degradation and full AIU/AHD95"); under torchrun every rank runs its own batch
(weak scaling, no data-path collective; the integer counts / HD values are all-gathered once per step).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 64] [--impl reference]

`--impl reference` times the reference's own algorithm on the host CPU cores (the oracle port under oracle/,
pinned to the unmodified reference by tests/golden) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SR+seg+AIU/AHD95 images/sec (448^2 HR, x4)"
UNIT = "images/s"
HR = 448
# algorithmic (reference-dense) forward FLOPs per 448^2 image, BASELINE.md section 3
DENSE_GFLOP_PER_IMG = 2255.38


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured"
    except Exception:
        return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active")
                                                          for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- reference arm
def cpu_reference_step(hr, mask, params, sd):
    """The reference's algorithm on the CPU (oracle port): degrade -> JointModel.forward -> AIU + HD sweep."""
    import torch
    from oracle import degrade_ref, metrics_ref, torch_ref
    with torch.no_grad():
        lr, _, _ = degrade_ref.degrade(hr, params)
        sr, seg, kp, _ = torch_ref.joint_forward(sd, lr)
    inter, union = metrics_ref.iou_counts(seg.numpy(), mask.numpy())
    hd, msd = metrics_ref.distance_metrics(seg.numpy(), mask.numpy(), 50)
    return metrics_ref.iou_from_counts(inter, union), hd, msd


def reference_available():
    from oracle import ref_harness as rh
    return rh.available()


_REF = {}


def real_reference_step(hr, mask, sd):
    """The UNMODIFIED reference on the CPU (baseline/_ref through oracle/ref_harness.py): `set_blur` + `conv_kernel2d` +
    `FactorResize` (crack_dataset.py:52-62), `JointModel.forward` (build_model.py:441-500), the threshold sweep + `IoU`
    (inference.py:49-53,111,119) and `calc_distance_metrics` (inference.py:293-336) -- the reference's own functions."""
    import contextlib
    import io
    import torch
    from oracle import ref_harness as rh
    if not _REF:
        rh.setup()
        from model.data.blur.blur import conv_kernel2d, set_blur
        from model.data.transforms.transforms import FactorResize
        IoU, calc, _ = rh.metric_fns()
        m = rh.joint_model(rh.make_cfg("PSPNet"))
        m.load_state_dict(sd, strict=True)
        _REF.update(model=m, set_blur=set_blur, conv=conv_kernel2d, resize=FactorResize(4, "bicubic"), iou=IoU(), calc=calc)
    R = _REF
    th = torch.Tensor([i * 0.01 for i in range(1, 100)]).view(99, 1, 1)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        lrs = []
        for i in range(hr.shape[0]):
            k = R["set_blur"](21, mode="gaus", isotropic=False).to("cpu")
            lrs.append(R["resize"](R["conv"](hr[i], k).to("cpu")))
        lr = torch.stack(lrs)
        sr, seg, kp = R["model"](lr, torch.zeros(len(lr), 1, 7, 7))
        bi = (seg - th > torch.Tensor([0])).float()
        iou = R["iou"](bi, mask)
        hd, msd, _, _ = R["calc"](bi, mask, 0, 0)
    return iou, hd, msd


def run_reference(args):
    import torch
    from csbsr_b200.utils import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.model_state_dict()
    n_img = 1                                   # bounded sample: one 448^2 image per step (~5-30 s of CPU work)
    hr, mask = synth.batch(0, n_img, HR)
    params = synth.degradation_params(n_img)
    real = reference_available()
    step = (lambda: real_reference_step(hr, mask, sd)) if real else (lambda: cpu_reference_step(hr, mask, params, sd))
    budget_s = 170.0
    t0 = time.time()
    step()                                      # one untimed pass (model construction, first-touch)
    t0 = time.time()
    step()
    per = max(time.time() - t0, 1e-3)
    steps = max(1, min(args.steps, int(budget_s / per)))
    t0 = time.time()
    for _ in range(steps):
        step()
    dt = time.time() - t0
    value = n_img * steps / dt
    what = ("the UNMODIFIED reference (baseline/_ref: set_blur + conv_kernel2d + FactorResize, JointModel.forward fp32 on torch CPU, "
            "IoU sweep + calc_distance_metrics)" if real else
            "the oracle port of the reference (degrade + KBPN + PSPNet fp32 on torch CPU, numpy/scipy AIU + HD sweep)")
    sample = "%d x 448^2 image(s) per step through %s, %d of %d requested steps run within the %.0f s budget" % (
        n_img, what, steps, args.steps, budget_s)
    kind = "reference" if real else "port"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": 2, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "CSBSR w/ PSPNet x4 eval, 448^2 HR, on-the-fly degradation, AIU+HD sweep (99 thr)",
                       "batch_per_step": n_img},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- our arm
def train_leg(args, c, dev, world, rank, timed):
    """Second half of BASELINE.json's metric ("train steps/sec"): config #3, the joint training step at iteration 40000
    (all phases active, w^F on with m^F = 1), 224^2 HR crops, 8 images per GPU: on-device degradation -> forward ->
    calc_loss -> backward (conv dgrad / wgrad on the tcgen05 engine) -> NCCL gradient all-reduce -> fused Adam."""
    import torch
    from csbsr_b200 import _lib
    from csbsr_b200.engine.optim import FusedAdam
    from csbsr_b200.engine.trainer import GraphedTrainStep, train_step
    from csbsr_b200.modeling.build_model import JointModelWithLoss
    from csbsr_b200.utils import synth
    tc = c.clone()
    tc.SOLVER.SEG_FAIL_ORIENTED_WEIGHT4SS_AMP = 1.0
    bt, size = 8, 224
    m = JointModelWithLoss(tc, num_train_ds=1000, resume_iter=40000)
    m.load_state_dict(synth.model_state_dict(), strict=True)
    m.to(dev).train()
    opt = FusedAdam(m.parameters(), lr=tc.SOLVER.LR)
    hr, mask = synth.batch(1000 + rank * bt, bt, size)
    hr, mask = hr.to(dev), mask.to(dev)
    params = torch.as_tensor(synth.degradation_params(bt, seed=50 + rank)).to(dev)
    it = [40000]

    graphed = GraphedTrainStep(m, opt, tc, world)

    state = {"eager": bool(args.no_graph)}

    def step():
        it[0] += 1
        if not state["eager"]:
            try:
                return graphed(it[0], hr, mask, params)[0]
            except Exception as e:                              # noqa: BLE001 -- capture not possible: launch eagerly
                if graphed.graphs:                              # a replay failed after a successful capture: a real error
                    raise
                state["eager"] = True
                torch.cuda.synchronize()
                opt.flat_g.zero_()
                if rank == 0:
                    print("training-step graph capture failed, launching eagerly: %r" % (e,), file=sys.stderr)
        return train_step(m, opt, tc, it[0], hr, mask, params, world)[0]

    for _ in range(3):
        step()
    l0 = _lib.LAUNCHES
    ms, loss = timed(step, args.train_steps)
    launches = graphed.launches_per_step if not state["eager"] else (_lib.LAUNCHES - l0) // args.train_steps
    ms /= args.train_steps
    dense_tf = 3 * DENSE_GFLOP_PER_IMG * (size / HR) ** 2 * bt / 1e3 / (ms * 1e-3)
    return {"metric": "CSBSR w/ PSPNet joint training steps/sec", "value": 1000.0 / ms, "unit": "steps/s",
            "images_per_sec": world * bt * 1000.0 / ms, "ms_per_step": ms, "n_gpus": world, "batch_per_gpu": bt, "crop": size,
            "steps": args.train_steps, "warmup": 3, "loss": float(loss.item()), "dtype": "bf16 activations / fp32 master weights",
            "gpu_launches": int(launches),
            "reference_dense_tflops_equiv_per_gpu": dense_tf,
            "config": "iteration 40000 (joint phase), SR L1 + pseudo-LR L1 + BoundaryCombo with w^F (m^F=1), Dropout2d and "
                      "BatchNorm batch statistics on, Adam lr 2e-5; gradients all-reduced over NCCL when n_gpus > 1; "
                      "everything between the convs on own kernels (csrc/glue.cu, cuDNN disabled); forward + loss + backward replayed "
                      "from a CUDA graph (eager launches if capture is unavailable), all-reduce and fused Adam launched per step"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--chunk", type=int, default=16, help="images per pass through the KBPN engine (measured: 8 -> 434, 16 -> 447, 32 -> 447 img/s)")
    ap.add_argument("--seg-chunk", type=int, default=0, help="images per pass through the segmentation engine (0 = the model's default)")
    ap.add_argument("--mchunk", type=int, default=16, help="images per csbsr_seg_metrics call")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step leg (BASELINE metric part 2)")
    ap.add_argument("--no-graph", action="store_true", help="launch the eval hot path and the training step eagerly instead of replaying CUDA graphs")
    ap.add_argument("--detector", default="PSPNet", choices=["PSPNet", "PSPNet_BlurSkip", "HRNet_OCR"],
                    help="segmentation net of the eval leg: PSPNet = the headline config (#2); the others are configs #5 / #4")
    ap.add_argument("--train-steps", type=int, default=5)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from csbsr_b200 import _lib, kernels as K
    from csbsr_b200.config import cfg
    from csbsr_b200.data import degrade as G
    from csbsr_b200.engine import distributed as D
    from csbsr_b200.engine import inference as E
    from csbsr_b200.modeling.build_model import JointModel
    from csbsr_b200.utils import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()                                   # fail loudly if the CUDA library is missing
    if not _lib.lib().csbsr_device_ok():
        raise _lib.CsbsrError("bench.py needs an sm_100 device")

    B = args.batch
    c = cfg.clone()
    c.merge_from_file(os.path.join(ROOT, "config", "config_csbsr_pspnet.yaml"))
    c.MODEL.DETECTOR_TYPE = args.detector
    model = JointModel(c)
    if args.detector == "PSPNet":
        model.load_state_dict(synth.model_state_dict(), strict=True)
    else:
        from csbsr_b200.modeling import params as MP
        sd_ = MP.synth_state_dict(MP.kbpn_param_shapes(), prefix="sr_model.")
        seg_shapes = MP.hrnet_ocr_param_shapes() if args.detector == "HRNet_OCR" else \
            MP.pspnet_param_shapes(blur_dim=c.BLUR.KERNEL_SIZE_OUTPUT ** 2)
        sd_.update(MP.synth_state_dict(seg_shapes, prefix="segmentation_model."))
        model.load_state_dict(sd_, strict=True)
        args.no_train = True                      # the training leg is the PSPNet config (#3)
    model.chunk = args.chunk
    if args.seg_chunk > 0:
        model.seg_chunk = args.seg_chunk

    # synthetic inputs: a few distinct images tiled to the batch (generation is untimed); each rank its own shard
    n_unique = min(B, 16)
    hr_u, mask_u = synth.batch(rank * B, n_unique, HR)
    reps = (B + n_unique - 1) // n_unique
    hr_host = hr_u.repeat(reps, 1, 1, 1)[:B].contiguous().pin_memory()
    mask_host = mask_u.repeat(reps, 1, 1, 1)[:B].contiguous().pin_memory()
    params = synth.degradation_params(B, seed=5 + rank)
    hr_dev, mask_dev = hr_host.to(dev), mask_host.to(dev)
    params_dev = torch.as_tensor(params).to(dev)
    mchunk = args.mchunk

    keep = {}

    def hot_path(hr, mask):
        """degrade -> SR -> seg -> metrics for one batch resident on the device; returns device tensors."""
        lr, _ = G.degrade(hr, params_dev)
        sr, seg, kp = model(lr, None)
        keep["seg"] = seg                       # parity self-check below reads the probability maps of the last eager call
        outs = []
        for i in range(0, B, mchunk):
            r = E.seg_metrics(seg[i:i + mchunk], mask[i:i + mchunk], with_hd=True, to_host=False)
            outs.append(D.pack_metrics(r["inter"], r["union"], r["hd"], r["msd"]))
        return torch.cat(outs, 0)               # [B, 4*99] fp64

    def gather(res):
        return D.gather_rows(res)                # the only exchange: B x 396 fp64 per rank (no-op at world size 1)

    # The hot path of one batch is ~1400 launches: replay it from a CUDA graph (same kernels, no per-launch CPU cost).
    # Static input buffers; falls back to eager launches if capture is not possible.
    graph_state = {"graph": None, "hr": hr_dev.clone(), "mask": mask_dev.clone(), "out": None}

    def try_capture():
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                hot_path(graph_state["hr"], graph_state["mask"])
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                graph_state["out"] = hot_path(graph_state["hr"], graph_state["mask"])
            graph_state["graph"] = g
        except Exception as e:                                  # noqa: BLE001
            graph_state["graph"] = None
            torch.cuda.synchronize()
            if rank == 0:
                print("cuda graph capture failed, launching eagerly: %r" % (e,), file=sys.stderr)

    def run_hot_path(hr, mask):
        if graph_state["graph"] is None:
            return hot_path(hr, mask)
        if hr.data_ptr() != graph_state["hr"].data_ptr():
            graph_state["hr"].copy_(hr, non_blocking=True)
            graph_state["mask"].copy_(mask, non_blocking=True)
        graph_state["graph"].replay()
        return graph_state["out"]

    def step_device():
        return gather(run_hot_path(graph_state["hr"], graph_state["mask"]))

    # end-to-end leg: every step's inputs come from pinned host memory through the repo's prefetcher (data/prefetch.py: the copy
    # of the next step's batch runs on its own stream under the current step, like a DataLoader with pin_memory feeding the
    # reference's loop), and every step's result is read back to the host
    from csbsr_b200.data.prefetch import DevicePrefetcher

    def host_batches():
        while True:
            yield (hr_host, mask_host)
    e2e_state = {"it": None}

    def step_e2e():
        if e2e_state["it"] is None:
            e2e_state["it"] = DevicePrefetcher(host_batches(), dev)
        hr, mask = next(e2e_state["it"])
        res = gather(run_hot_path(hr, mask)).cpu().numpy()
        inter, union, hd = res[:, :99], res[:, 99:198], res[:, 198:297]
        iou = (inter + 1e-5) / (union + 1e-5)
        return float(np.mean(iou)), float(np.mean(hd))     # AIU, AHD (inference.py:171-173)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    first = hot_path(hr_dev, mask_dev)                          # first call packs the weights and sizes the workspaces
    torch.cuda.synchronize()
    # parity self-check (oracle as the checker only): the step's integer I / U counts of every image and the HD / MSD
    # rows of image 0 must equal the reference's sweep (inference.py:111-121, 293-336) on the same probability maps
    parity = None
    if rank == 0:
        from oracle import metrics_ref
        seg_h, mask_h = keep["seg"].cpu().numpy(), mask_dev.cpu().numpy()
        inter_o, union_o = metrics_ref.iou_counts(seg_h, mask_h)
        hd_o, msd_o = metrics_ref.distance_metrics(seg_h[:1], mask_h[:1], 50)
        inter_d, union_d, hd_d, msd_d = D.unpack_metrics(first)
        parity = bool(np.array_equal(inter_d, inter_o) and np.array_equal(union_d, union_o)
                      and np.array_equal(hd_d[:1], hd_o) and np.array_equal(msd_d[:1], msd_o))
        if not parity:
            raise AssertionError("bench.py: AIU counts / HD / MSD of the step differ from the oracle sweep")
    l_cap = _lib.LAUNCHES
    if not args.no_graph:
        try_capture()
    launches_per_step = (_lib.LAUNCHES - l_cap) // 2 if graph_state["graph"] is not None else None
    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.LAUNCHES
    ms_dev, _ = timed(step_device, args.steps)
    launches = launches_per_step if launches_per_step is not None else (_lib.LAUNCHES - l0) // max(args.steps, 1)
    step_e2e()                                   # warm-up of the e2e path (allocates the prefetcher's device buffers)
    torch.cuda.synchronize()
    e2e_state["it"] = None                       # the timed region starts cold: its first step issues its own H2D copy, so the K timed
    ms_e2e, (aiu, ahd) = timed(step_e2e, args.steps)   # steps contain K + 1 copies (the last prefetch is never consumed)
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel (conv_igemm_kernel), measured live with CUDA events per launch
    K.PROFILE = []
    hot_path(hr_dev, mask_dev)                  # eager launches: the per-launch events cannot be recorded inside a graph replay
    torch.cuda.synchronize()
    conv_ms = sum(r[2].elapsed_time(r[3]) for r in K.PROFILE)
    conv_useful = sum(r[4] for r in K.PROFILE)
    conv_padded = sum(r[1] for r in K.PROFILE)
    n_conv = len(K.PROFILE)
    by_layer = {}
    for r in K.PROFILE:                       # (label, padded flops, event, event, useful flops, bytes)
        a = by_layer.setdefault(r[0], [0, 0.0, 0.0])
        a[0] += 1; a[1] += r[2].elapsed_time(r[3]); a[2] += r[4]
    top_layers = [{"layer": k, "launches": v[0], "ms": round(v[1], 3), "tflops": round(v[2] / max(v[1], 1e-9) / 1e9, 1)}
                  for k, v in sorted(by_layer.items(), key=lambda kv: -kv[1][1])[:16]]
    K.PROFILE = None
    peak_tf, peak_bw, peak_src = _peaks()
    traffic = None                       # DRAM bytes of the conv kernel per step, from the committed ncu capture
    try:
        with open(os.path.join(ROOT, "profiles", "r02_conv_traffic.json")) as f:
            traffic = json.load(f)["dram_bytes_per_image"] * B
    except Exception:
        pass
    achieved_tf = conv_useful / (conv_ms * 1e-3) / 1e12
    value = world * B * args.steps / (ms_dev * 1e-3)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
    h2d = int(hr_host.numel() * 4 + mask_host.numel() * 4)
    d2h = int(world * B * 4 * 99 * 8)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "CSBSR w/ " + args.detector + " x4 eval (config_csbsr_pspnet.yaml), batch %d x 448^2 HR per GPU, on-the-fly "
                               "anisotropic-blur degradation, AIU + HD(p50)/MSD sweep over 99 thresholds" % B,
                   "batch_per_gpu": B, "chunk": args.chunk, "seg_chunk": model.seg_chunk, "metrics_chunk": mchunk, "hr": HR, "scale": 4, "weights": "synthetic random-init (seed 1121)",
                   "l2": "per-step inputs (%.0f MB) and activations exceed the 126 MB L2; no flush needed" % (h2d / 1e6),
                   "aiu": aiu, "ahd_p50": ahd,
                   "parity_check": "I/U counts of all %d images and HD/MSD of image 0 bit-equal to the oracle sweep: %s" % (B, parity),
                   "launch": "CUDA graph replay of the hot path" if graph_state["graph"] is not None else "eager launches"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "csbsr::conv_igemm_kernel", "achieved": achieved_tf, "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": achieved_tf / peak_tf, "traffic": traffic,
                     "traffic_note": "dram__bytes_read+write of all conv launches of one step (profiles/r02_launches.md), bytes",
                     "peak_source": peak_src + " (sustained bf16)",
                     "launches_per_step": n_conv, "kernel_ms_per_step": conv_ms, "share_of_step": conv_ms / (ms_dev / args.steps),
                     "useful_gflop_per_step": conv_useful / 1e9, "padded_gflop_per_step": conv_padded / 1e9,
                     "reference_dense_gflop_per_step": DENSE_GFLOP_PER_IMG * B,
                     "reference_dense_tflops_equiv": DENSE_GFLOP_PER_IMG * B / 1e3 / (ms_dev / args.steps * 1e-3),
                     "top_layers": top_layers},
        "clocks": clocks,
    }
    if rank == 0 and world == 1:
        # HBM-class kernels against their algorithmic bytes (north_star: degradation / metric kernels vs the HBM peak)
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import bench_hbm_kernels
        line["roofline_hbm"] = bench_hbm_kernels.run(B)
        torch.cuda.empty_cache()
    if not args.no_train:
        line["train"] = train_leg(args, c, dev, world, rank, timed)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            sd = synth.model_state_dict()
            real = reference_available()
            cstep = (lambda: real_reference_step(hr_u[:1], mask_u[:1], sd)) if real else \
                (lambda: cpu_reference_step(hr_u[:1], mask_u[:1], params[:1], sd))
            t0 = time.time()
            cstep()
            dt = time.time() - t0
            line["cpu_baseline"] = {"value": 1.0 / dt, "unit": UNIT, "cores": cores, "kind": "reference" if real else "port",
                                    "sample": "1 x 448^2 image through %s (degrade + KBPN + PSPNet fp32 torch CPU + AIU/HD "
                                              "sweep), single run incl. model construction, %.1f s"
                                              % ("the unmodified reference (baseline/_ref)" if real else "the oracle port", dt)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
