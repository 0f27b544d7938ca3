"""PNG writers of the evaluation loop with the reference's paths and value conventions (model/utils/save_output.py:16-73;
called from model/engine/inference.py:101-118 when --sf_save_image is set).  Host-side I/O: tensors are read back once."""
import os

import numpy as np
import torch


def _to_pil(t):
    """torchvision ToPILImage on a float tensor [C, H, W] in [0, 1]: mul(255) and truncate to uint8."""
    from PIL import Image
    a = t.detach().float().cpu().mul(255).byte().numpy()
    if a.shape[0] == 1:
        return Image.fromarray(a[0], mode="L")
    return Image.fromarray(np.transpose(a, (1, 2, 0)), mode="RGB")


def save_img(dirname, sr_preds, fname):
    os.makedirs(os.path.join(dirname, "images"), exist_ok=True)
    for i in range(sr_preds.shape[0]):
        _to_pil(sr_preds[i]).save(os.path.join(dirname, "images", "%s" % fname[i]))


def save_mask(args, segment_preds, fname, iou_th, add_path=""):
    d = os.path.join(args.output_dirname, "masks%s" % add_path, "th_%.2f" % iou_th)
    os.makedirs(d, exist_ok=True)
    for i in range(segment_preds.shape[0]):
        _to_pil(segment_preds[i]).save(os.path.join(d, "%s" % fname[i]))


def save_kernel(args, kernel_preds, fname, num_batch, add_path=""):
    num_patch = kernel_preds.shape[0] // num_batch
    d0 = os.path.join(args.output_dirname, "kernels%s" % add_path)
    d1 = os.path.join(args.output_dirname, "kernels%s_origin" % add_path)
    os.makedirs(d0, exist_ok=True)
    os.makedirs(d1, exist_ok=True)
    for i in range(num_batch):
        stem = ("%s" % fname[i]).replace(".png", "")
        for j in range(num_patch):
            k = kernel_preds[i * num_patch + j]
            _to_pil(k / torch.max(k)).save(os.path.join(d0, "%s_%d.png" % (stem, j)))
            _to_pil(k / torch.sum(k)).save(os.path.join(d1, "%s_%d_origin.png" % (stem, j)))
