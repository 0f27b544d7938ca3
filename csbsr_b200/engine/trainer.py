"""do_train (reference model/engine/trainer.py:57-72, calc_loss :406-438) for one process per GPU.

A step is: on-device degradation of the HR crop batch (crack_dataset.py:51-62) -> JointModelWithLoss.forward ->
calc_loss -> backward -> NCCL SUM all-reduce of the flat gradient (replaces nn.DataParallel, train.py:117-121) ->
fused Adam (+ gradient zeroing) -> LambdaLR step.  Per-rank batches are the reference's per-GPU chunks, so BatchNorm
statistics and the (B,B,H,W) w^F broadcasting stay per replica exactly as under DataParallel."""
import time

import torch

from ..data import degrade as G
from .losses import calc_loss


def train_step(model, optimizer, cfg, iteration, hr, mask, params, world_size=1):
    """One optimisation step on this rank's shard. `params` = (theta, sigma_x, sigma_y) float64 [B,3] for the blur."""
    lr_img, kernels = G.degrade(hr, params, ksize=cfg.BLUR.KERNEL_SIZE_OUTPUT, factor=cfg.MODEL.SCALE_FACTOR)
    seg_loss, sr_loss, seg, sr, kp = model(iteration, lr_img, sr_targets=hr, segment_targets=mask,
                                           kernel_targets=kernels.unsqueeze(1))
    loss = calc_loss(sr_loss, seg_loss.mean(), cfg.SOLVER.TASK_LOSS_WEIGHT, iteration, cfg)
    loss.backward()
    if world_size > 1:
        optimizer.all_reduce_grads()
    optimizer.step(world_size)
    optimizer.scheduler_step()
    return loss.detach(), seg_loss.detach().mean(), sr_loss.detach().mean()


def do_train(args, cfg, model, optimizer, batches, rank=0, world_size=1, log=print):
    """`batches` yields (iteration, hr [B,3,H,W], mask [B,1,H,W], blur params [B,3]) for THIS rank."""
    model.train()
    t0 = time.time()
    for iteration, hr, mask, params in batches:
        # boundary-loss alpha schedule, poked from the trainer before every step (fix_1st_stage_model_params,
        # trainer.py:495-508): frozen at its start value during SR pre-training, one schedule tick per iteration after it
        if cfg.SOLVER.SR_PRETRAIN_ITER[0] <= iteration < cfg.SOLVER.SR_PRETRAIN_ITER[1]:
            model.ss_loss_fn.fix_alpha, model.ss_loss_fn.iter = True, 1
        else:
            model.ss_loss_fn.fix_alpha = False
            model.ss_loss_fn.update_alpha()
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
