#!/usr/bin/env python
"""Evaluation entry point with the reference's command line (reference test.py:83-157):

    python test.py <test_dir> <iter_or_weight_name> [--batch_size 12] [--num_gpus 1] [--test_surface_distance]
                   [--config_file FILE] [--trained_model FILE] [--test_blured_name NAME] [--output_dirname DIR]

<test_dir>/config.yaml and <test_dir>/model/<name>.pth are used unless overridden (test.py:105-117).  Offline (no
dataset, no checkpoint): `--synthetic N` evaluates N seeded synthetic 448x448 crack images and, when the checkpoint
file is missing, loads the deterministic synthetic weights.  Multi-GPU: launch with torchrun, one process per GPU;
each rank evaluates a contiguous shard of the test set and the per-image metrics are all-gathered."""
import argparse
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser(description="Crack Segmentation with Blind Super Resolution (CSBSR) -- B200 build")
    ap.add_argument("test_dir", type=str)
    ap.add_argument("iter_or_weight_name", type=str)
    ap.add_argument("--output_dirname", type=str, default=None)
    ap.add_argument("--config_file", type=str, default=None, metavar="FILE")
    ap.add_argument("--test_blured_name", type=str, default=None)
    ap.add_argument("--num_workers", type=int, default=0)
    ap.add_argument("--batch_size", type=int, default=12)
    ap.add_argument("--num_gpus", type=int, default=1)
    ap.add_argument("--test_surface_distance", action="store_true")
    ap.add_argument("--trained_model", type=str, default=None)
    ap.add_argument("--hd_percentile", type=float, default=50.0, help="the reference ships 50 (inference.py:302)")
    ap.add_argument("--synthetic", type=int, default=0, help="evaluate N synthetic images instead of the dataset")
    args = ap.parse_args()

    from csbsr_b200.config import cfg
    from csbsr_b200.data import crack_dataset as DS
    from csbsr_b200.engine import distributed as D
    from csbsr_b200.engine.inference import inference_for_ss
    from csbsr_b200.modeling.build_model import JointModel
    from csbsr_b200.utils import synth

    name = args.iter_or_weight_name
    is_iter = not re.search(r"[^0-9]", name)
    out_dir = ("iter_%s" % name) if is_iter else name
    model_fname = ("iteration_%s" % name) if is_iter else name
    td = args.test_dir if args.test_dir.endswith("/") else args.test_dir + "/"
    config_file = args.config_file or td + "config.yaml"
    trained_model = args.trained_model or td + "model/%s.pth" % model_fname
    output_dirname = args.output_dirname or td + "eval_AIU/%s" % out_dir

    img_size = cfg.INPUT.IMAGE_SIZE                    # captured BEFORE the merge, like the reference (test.py:119-120)
    if os.path.exists(config_file):
        print("Configration file is loaded from {}".format(config_file))
        cfg.merge_from_file(config_file)
    elif not args.synthetic:
        raise FileNotFoundError(config_file)
    else:
        cfg.merge_from_file(os.path.join(ROOT, "config", "config_csbsr_pspnet.yaml"))
    if args.test_blured_name is not None:
        cfg.DATASET.TEST_BLURED_NAME = args.test_blured_name
    cfg.OUTPUT_DIR = output_dirname
    cfg.INPUT.IMAGE_SIZE = img_size
    cfg.freeze()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))

    model = JointModel(cfg)
    if os.path.exists(trained_model):
        sd = torch.load(trained_model, map_location="cpu")
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}      # fix_model_state_dict, misc.py:35-44
        model.load_state_dict(sd, strict=True)
        print("Trained model is loaded from {}".format(trained_model))
    elif args.synthetic:
        sd = synth.model_state_dict()
        if cfg.MODEL.DETECTOR_TYPE in ("PSPNet_BlurSkip", "HRNet_OCR"):
            from csbsr_b200.modeling import params as P
            sd = P.synth_state_dict(P.kbpn_param_shapes(), prefix="sr_model.")
            seg_shapes = P.hrnet_ocr_param_shapes() if cfg.MODEL.DETECTOR_TYPE == "HRNet_OCR" else \
                P.pspnet_param_shapes(blur_dim=cfg.BLUR.KERNEL_SIZE_OUTPUT ** 2)
            sd.update(P.synth_state_dict(seg_shapes, prefix="segmentation_model."))
        model.load_state_dict(sd, strict=True)
        print("checkpoint %s not found: synthetic weights loaded" % trained_model)
    else:
        raise FileNotFoundError(trained_model)
    model.eval()

    full = DS.SyntheticCrackTestSet(args.synthetic, img_size[0]) if args.synthetic else DS.CrackDataSetTest(cfg)
    lo, hi = D.shard_range(len(full))
    shard = torch.utils.data.Subset(full, range(lo, hi))
    loader = torch.utils.data.DataLoader(shard, batch_size=args.batch_size, shuffle=False, num_workers=args.num_workers,
                                         collate_fn=DS.collate)
    with torch.no_grad():
        inference_for_ss(model, loader, test_surface_distance=args.test_surface_distance, percent=args.hd_percentile,
                         output_dir=output_dirname)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
