#!/usr/bin/env python
"""Training entry point with the reference's command line (reference train.py:127-170):

    python train.py --config_file config/config_csbsr_pspnet.yaml [--output_dirname DIR] [--num_gpus N]
                    [--resume_iter I] [--log_step 50] [--save_step 2000] [--eval_step 2000] [--num_workers 2]

Multi-GPU: launch with torchrun, one process per GPU (the reference's nn.DataParallel is replaced by per-rank batch
shards + an NCCL all-reduce of the gradients).  Offline (`--synthetic N`): trains on N seeded synthetic crack images
with on-device degradation.  All training phases are covered (the three SR pre-training phases and the joint phase,
see JointModelWithLoss.apply_phase).  SOLVER.BATCH_SIZE is the GLOBAL batch, split over the ranks like nn.DataParallel
splits it over its GPUs (reference train.py:108-121)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser(description="Crack Segmentation with Blind Super Resolution (CSBSR) -- B200 build")
    ap.add_argument("--config_file", type=str, default="./config/config_csbsr_pspnet.yaml", metavar="FILE")
    ap.add_argument("--output_dirname", type=str, default="")
    ap.add_argument("--num_workers", type=int, default=2)
    ap.add_argument("--log_step", type=int, default=50)
    ap.add_argument("--save_step", type=int, default=2000)
    ap.add_argument("--eval_step", type=int, default=2000)
    ap.add_argument("--num_gpus", type=int, default=1)
    ap.add_argument("--resume_iter", type=int, default=0)
    ap.add_argument("--max_iter", type=int, default=None, help="stop after this iteration (default SOLVER.MAX_ITER)")
    ap.add_argument("--synthetic", type=int, default=0, help="train on N synthetic crack images")
    ap.add_argument("--crop", type=int, default=None, help="HR crop size (default INPUT.IMAGE_SIZE of the config)")
    ap.add_argument("--init_synthetic", action="store_true",
                    help="offline demo: start the schedules at --resume_iter on seeded synthetic weights instead of a checkpoint "
                         "(explicit only -- a missing checkpoint is an error otherwise)")
    ap.add_argument("--cuda_graph", type=int, default=1, help="replay forward+loss+backward from a CUDA graph (1) or launch eagerly (0)")
    args = ap.parse_args()

    from csbsr_b200.config import cfg
    from csbsr_b200.engine.optim import FusedAdam, UpDownScheduler
    from csbsr_b200.engine.trainer import do_train
    from csbsr_b200.modeling import params as P
    from csbsr_b200.modeling.build_model import JointModelWithLoss
    from csbsr_b200.utils import synth

    cfg.merge_from_file(args.config_file)
    if args.output_dirname:
        cfg.OUTPUT_DIR = args.output_dirname
    cfg.freeze()
    if int(os.environ.get("RANK", "0")) == 0 and (args.resume_iter == 0 or args.init_synthetic):
        # the run's config travels with its output so that `test.py <output_dir> <iter>` finds it (reference train.py:150-153)
        import shutil
        os.makedirs(cfg.OUTPUT_DIR, exist_ok=True)
        shutil.copy2(args.config_file, os.path.join(cfg.OUTPUT_DIR, "config.yaml"))

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(cfg.SEED)

    dataset = None
    if not args.synthetic:
        from csbsr_b200.data.crack_dataset import CrackDataSet
        dataset = CrackDataSet(cfg, seed=cfg.SEED + rank)
        n_train = int(len(dataset) * cfg.SOLVER.TRAIN_DATASET_RATIO)     # train.py:52 (the rest is the validation split)
        if n_train == 0:
            raise FileNotFoundError("no *.jpg images under %s (use --synthetic N to train offline)" % dataset.image_dir)
    model = JointModelWithLoss(cfg, num_train_ds=args.synthetic or n_train, resume_iter=args.resume_iter)
    ckpt = os.path.join(cfg.OUTPUT_DIR, "model", "iteration_{}.pth".format(args.resume_iter))
    if args.resume_iter > 0 and not args.init_synthetic:
        if not os.path.exists(ckpt):                       # the reference fails in torch.load here; never resume on random weights
            raise FileNotFoundError("--resume_iter %d: checkpoint %s does not exist" % (args.resume_iter, ckpt))
        sd = torch.load(ckpt, map_location="cpu")
        model.load_state_dict({(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}, strict=False)
        print("Resume from {}".format(ckpt))
    else:
        sd = P.synth_state_dict(P.kbpn_param_shapes(), prefix="sr_model.")
        blur_dim = cfg.BLUR.KERNEL_SIZE_OUTPUT ** 2 if cfg.MODEL.DETECTOR_TYPE == "PSPNet_BlurSkip" else None
        seg_shapes = (P.hrnet_ocr_param_shapes() if cfg.MODEL.DETECTOR_TYPE == "HRNet_OCR"
                      else P.pspnet_param_shapes(cfg.MODEL.NUM_CLASSES, blur_dim=blur_dim))
        sd.update(P.synth_state_dict(seg_shapes, prefix="segmentation_model."))
        model.load_state_dict(sd, strict=True)
    model.cuda()
    sched = UpDownScheduler(cfg.SOLVER.SR_PRETRAIN_ITER[1], args.resume_iter, cfg.SOLVER.SCHEDULER)
    optimizer = FusedAdam(model.parameters(), lr=cfg.SOLVER.LR, betas=(0.9, 0.999), eps=1e-8, lr_lambda=sched)

    size = args.crop or cfg.INPUT.IMAGE_SIZE[0]
    # nn.DataParallel splits SOLVER.BATCH_SIZE over the GPUs (train.py:108-121): each rank takes its chunk, so the
    # global batch and the alpha / LR schedules (per_epoch = n_train // BATCH_SIZE + 1) are the reference's
    if cfg.SOLVER.BATCH_SIZE % world != 0:
        raise ValueError("SOLVER.BATCH_SIZE=%d is not divisible by the %d ranks" % (cfg.SOLVER.BATCH_SIZE, world))
    per_rank = cfg.SOLVER.BATCH_SIZE // world
    if world > 1 and cfg.SOLVER.SYNC_BATCHNORM and rank == 0:
        # the reference converts to synchronised BatchNorm under DataParallel (train.py:98-100, statistics over the whole
        # batch); here every rank normalises with the statistics of its own chunk, like DataParallel without the conversion
        print("warning: SOLVER.SYNC_BATCHNORM is set, but BatchNorm statistics stay per rank (%d images each); running "
              "statistics are checkpointed from rank 0" % per_rank, file=sys.stderr)
    max_iter = args.max_iter or cfg.SOLVER.MAX_ITER
    rng = np.random.default_rng(cfg.SEED + rank)

    def batches():
        for it in range(args.resume_iter + 1, max_iter + 1):
            if dataset is not None:
                idx = rng.integers(0, n_train, size=per_rank)
                hr, mask = zip(*(dataset[int(i)] for i in idx))
            else:
                idx = rng.integers(0, args.synthetic, size=per_rank)
                hr, mask = zip(*(synth.crack_image(int(i), size) for i in idx))
            theta = rng.uniform(0, 180, per_rank) * np.pi / 180.0
            sig = rng.uniform(0.2, 4.0, (per_rank, 2))
            params = np.concatenate([theta[:, None], sig], axis=1)
            yield it, torch.stack(hr).cuda(), torch.stack(mask).cuda(), params

    do_train(args, cfg, model, optimizer, batches(), rank, world)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
