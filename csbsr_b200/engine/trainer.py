"""do_train (reference model/engine/trainer.py:57-72, calc_loss :406-438) for one process per GPU.

A step is: on-device degradation of the HR crop batch (crack_dataset.py:51-62) -> JointModelWithLoss.forward ->
calc_loss -> backward -> NCCL SUM all-reduce of the flat gradient (replaces nn.DataParallel, train.py:117-121) ->
fused Adam (+ gradient zeroing) -> LambdaLR step.  Per-rank batches are the reference's per-GPU chunks, so BatchNorm
statistics and the (B,B,H,W) w^F broadcasting stay per replica exactly as under DataParallel."""
import os
import time

import torch

from ..data import degrade as G
from .losses import calc_loss


def train_step(model, optimizer, cfg, iteration, hr, mask, params, world_size=1):
    """One optimisation step on this rank's shard. `params` = (theta, sigma_x, sigma_y) float64 [B,3] for the blur."""
    ar = _overlap(model, optimizer) if world_size > 1 and AR_OVERLAP else None   # installs the hook the forward registers
    lr_img, kernels = G.degrade(hr, params, ksize=cfg.BLUR.KERNEL_SIZE_OUTPUT, factor=cfg.MODEL.SCALE_FACTOR)
    seg_loss, sr_loss, seg, sr, kp = model(iteration, lr_img, sr_targets=hr, segment_targets=mask,
                                           kernel_targets=kernels.unsqueeze(1))
    loss = calc_loss(sr_loss, seg_loss.mean(), cfg.SOLVER.TASK_LOSS_WEIGHT, iteration, cfg)
    loss.backward()
    if ar is not None:
        ar.finish()
    elif world_size > 1:
        optimizer.all_reduce_grads()
    optimizer.step(world_size)
    optimizer.scheduler_step()
    return loss.detach(), seg_loss.detach().mean(), sr_loss.detach().mean()


# CSBSR_AR_OVERLAP=1 starts the all-reduce of the segmentation net's gradient slice from an autograd hook, under the backward of
# the SR net, and captures the whole exchange in the step's CUDA graph (engine/distributed.py::OverlappedGradAllReduce; the
# host logic is covered by the gloo test).  Off by default: the persistent conv kernels fill every SM's register file, so NCCL's
# CTAs only run in the gaps between kernels, and the one 2-GPU verification run of this round did not finish within its time
# limit -- the default stays the exchange measured at N = 2, 4, 8 (bucketed all-reduce after the replay, 1.15 ms at N = 8).
AR_OVERLAP = os.environ.get("CSBSR_AR_OVERLAP") == "1"


def _overlap(model, optimizer):
    """The step's gradient exchange (engine/distributed.py::OverlappedGradAllReduce), created once per (model, optimizer)
    and hooked to the end of the segmentation net's backward."""
    from .distributed import OverlappedGradAllReduce
    ar = getattr(optimizer, "_overlap_ar", None)
    if ar is None:
        span = OverlappedGradAllReduce.prefix_span(model.named_parameters(), optimizer.params, optimizer.slots)
        ar = optimizer._overlap_ar = OverlappedGradAllReduce(optimizer.flat_g, span, optimizer.bucket_elems)
    model.on_seg_backward_done = ar.seg_done
    return ar


class GraphedTrainStep:
    """train_step with forward + loss + backward replayed from a CUDA graph (the step launches ~4000 small kernels; from
    Python it is launch-bound at ~60 ms, replayed it takes the ~50 ms of GPU time).  One graph per (training phase, input shape), re-captured when the alpha
    of the boundary loss changes (once per epoch; the stale graph and its memory pool are released first): the iteration
    only enters the forward through the phase switches, alpha is a kernel scalar.  Inputs are copied into static buffers; the gradient all-reduce and the fused Adam step stay outside the graph
    (the bias corrections change every step)."""

    def __init__(self, model, optimizer, cfg, world_size=1):
        self.model, self.opt, self.cfg, self.world = model, optimizer, cfg, world_size
        self.graphs = {}
        self.launches_per_step = 0

    def _phase_key(self, it):
        s = self.cfg.SOLVER
        inside = lambda r: r[0] <= it < r[1]
        return (inside(s.SR_PRETRAIN_ITER), inside(s.SR_SR_MODULE_PRETRAIN_ITER), inside(s.SR_KERNEL_MODULE_PRETRAIN_ITER),
                s.SR_KERNEL_MODULE_PRETRAIN_ITER[0] <= it < s.SR_KERNEL_MODULE_PRETRAIN_ITER[1] - 1,
                s.ORIENTED_WEIGHT_ITER <= it)

    def _body(self, st, it):
        cfg = self.cfg
        if self.world > 1 and AR_OVERLAP:
            st["ar"] = _overlap(self.model, self.opt)
        lr_img, kernels = G.degrade(st["hr"], st["params"], ksize=cfg.BLUR.KERNEL_SIZE_OUTPUT, factor=cfg.MODEL.SCALE_FACTOR)
        seg_loss, sr_loss, seg, sr, kp = self.model(it, lr_img, sr_targets=st["hr"], segment_targets=st["mask"],
                                                    kernel_targets=kernels.unsqueeze(1))
        seg_mean = seg_loss.mean()
        loss = calc_loss(sr_loss, seg_mean, cfg.SOLVER.TASK_LOSS_WEIGHT, it, cfg)
        loss.backward()
        if self.world > 1 and AR_OVERLAP:
            # opt-in: the exchange becomes part of the captured step -- the segmentation slice leaves from the autograd hook
            # (under the SR net's backward, on NCCL's stream), the rest here; a replay contains both
            st["ar"].finish()
        return loss.detach(), seg_mean.detach(), sr_loss.detach().mean()

    def _capture(self, it, hr, mask, params):
        st = {"hr": hr.clone(), "mask": mask.clone(), "params": torch.as_tensor(params).to(hr.device).clone()}
        buffers = [(b, b.clone()) for b in self.model.buffers()]         # warm-up must not advance the BN running stats
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._body(st, it)
        torch.cuda.current_stream().wait_stream(side)
        for b, saved in buffers:
            b.copy_(saved)
        self.opt.flat_g.zero_()
        from .. import _lib, kernels as K
        K.flush_pack_table()                 # job table of the multi-tensor weight pack on the device before the capture
        K.invalidate_packed()                # the captured step must contain the pack launch: replays follow optimizer steps
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.LAUNCHES
        with torch.cuda.graph(graph):
            st["out"] = self._body(st, it)
        self.launches_per_step = _lib.LAUNCHES - n0 + 1          # own kernels inside the graph + the Adam launch
        self.opt.flat_g.zero_()
        for b, saved in buffers:
            b.copy_(saved)
        st["graph"] = graph
        return st

    def __call__(self, iteration, hr, mask, params):
        # alpha of the boundary loss is a kernel scalar baked into the captured launches; it only ever decreases (one
        # tick per epoch), so a graph captured for an older alpha is never replayed again: drop it (and its private
        # memory pool, a full set of forward/backward activations) before capturing the next one.
        key = (self._phase_key(iteration), tuple(hr.shape))
        alpha = float(self.model.ss_loss_fn.alpha)
        st = self.graphs.get(key)
        if st is not None and st["alpha"] != alpha:
            del self.graphs[key]
            st.clear()
            st = None
            torch.cuda.empty_cache()
        if st is None:
            st = self.graphs[key] = self._capture(iteration, hr, mask, params)
            st["alpha"] = alpha
        st["hr"].copy_(hr, non_blocking=True)
        st["mask"].copy_(mask, non_blocking=True)
        st["params"].copy_(torch.as_tensor(params).to(st["params"].device), non_blocking=True)
        st["graph"].replay()                   # forward + loss + backward (+ the overlapped gradient all-reduce when opted in)
        if self.world > 1 and not AR_OVERLAP:
            self.opt.all_reduce_grads()
        self.opt.step(self.world)
        self.opt.scheduler_step()
        return st["out"]


def do_train(args, cfg, model, optimizer, batches, rank=0, world_size=1, log=print):
    """`batches` yields (iteration, hr [B,3,H,W], mask [B,1,H,W], blur params [B,3]) for THIS rank."""
    model.train()
    t0 = time.time()
    graphed = GraphedTrainStep(model, optimizer, cfg, world_size) if getattr(args, "cuda_graph", False) else None
    for iteration, hr, mask, params in batches:
        # boundary-loss alpha schedule, poked from the trainer before every step (fix_1st_stage_model_params,
        # trainer.py:495-508): frozen at its start value during SR pre-training, one schedule tick per iteration after it
        if cfg.SOLVER.SR_PRETRAIN_ITER[0] <= iteration < cfg.SOLVER.SR_PRETRAIN_ITER[1]:
            model.ss_loss_fn.fix_alpha, model.ss_loss_fn.iter = True, 1
        else:
            model.ss_loss_fn.fix_alpha = False
            model.ss_loss_fn.update_alpha()
        if graphed is not None:
            loss, seg_l, sr_l = graphed(iteration, hr, mask, params)
        else:
            loss, seg_l, sr_l = train_step(model, optimizer, cfg, iteration, hr, mask, params, world_size)
        if iteration % args.log_step == 0 and rank == 0:
            torch.cuda.synchronize()
            log("===> Iter: {:07d}, LR: {:.06f}, Cost: {:.2f}s, Loss: {:.6f} (seg {:.6f}, sr {:.6f}), alpha {:.3f}".format(
                iteration, optimizer.lr, time.time() - t0, loss.item(), seg_l.item(), sr_l.item(), model.ss_loss_fn.alpha))
            t0 = time.time()
        if args.save_step > 0 and iteration % args.save_step == 0 and rank == 0:
            import os
            os.makedirs(os.path.join(cfg.OUTPUT_DIR, "model"), exist_ok=True)
            torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()},
                       os.path.join(cfg.OUTPUT_DIR, "model", "iteration_{}.pth".format(iteration)))
            os.makedirs(os.path.join(cfg.OUTPUT_DIR, "optimizer"), exist_ok=True)      # reference trainer.py:119-129
            torch.save(optimizer.state_dict(), os.path.join(cfg.OUTPUT_DIR, "optimizer", "iteration_{}.pth".format(iteration)))
