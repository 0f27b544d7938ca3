"""Test-time data sources with the reference's on-disk convention (model/data/crack_dataset.py:71-142) and a
synthetic stand-in for offline runs.  Both yield per item: (lr[3,h,w], hr[3,H,W], mask[1,H,W], kernel[1,k,k], name).

Directory layout of the reference (README.md:98-107, crack_dataset.py:73-77):
    <TEST_IMAGE_DIR>/*.jpg, <TEST_MASK_DIR>/<same name>.jpg,
    <TEST_BLURED_DIR>/<TEST_BLURED_NAME>/kernels/<name>.png, .../lr_images/<name>.png
Images are converted to [0,1] floats as the reference's ConvertFromInts + ToTensor transforms do; one 112x112 LR
patch per 448x448 image (the reference's SplitPatch is the identity at this size)."""
import os
from pathlib import Path

import numpy as np
import torch


class CrackDataSetTest(torch.utils.data.Dataset):
    def __init__(self, cfg):
        from PIL import Image  # noqa: F401  (fail early when PIL is unavailable)
        self.image_dir = cfg.DATASET.TEST_IMAGE_DIR
        self.mask_dir = cfg.DATASET.TEST_MASK_DIR
        base = os.path.join(cfg.DATASET.TEST_BLURED_DIR, cfg.DATASET.TEST_BLURED_NAME)
        self.kernel_dir, self.lr_dir = os.path.join(base, "kernels"), os.path.join(base, "lr_images")
        self.fnames = sorted(p.name for p in Path(self.image_dir).glob("*.jpg"))
        if not self.fnames:
            raise FileNotFoundError("no *.jpg under %s" % self.image_dir)

    def __len__(self):
        return len(self.fnames)

    @staticmethod
    def _load(path, gray=False):
        from PIL import Image
        a = np.asarray(Image.open(path), dtype=np.float32) / 255.0
        if a.ndim == 2:
            a = a[:, :, None]
        return torch.from_numpy(a).permute(2, 0, 1).contiguous()

    def __getitem__(self, i):
        name = self.fnames[i]
        hr = self._load(os.path.join(self.image_dir, name))
        mask = self._load(os.path.join(self.mask_dir, name))[:1]
        png = name.replace("jpg", "png")
        kernel = self._load(os.path.join(self.kernel_dir, png))[:1]
        kernel = kernel / kernel.sum()
        lr = self._load(os.path.join(self.lr_dir, png))
        return lr, hr, mask, kernel, name


class CrackDataSet(torch.utils.data.Dataset):
    """Training set reader (reference CrackDataSet, model/data/crack_dataset.py:28-69, with the TrainTransforms of
    config_csbsr_pspnet.yaml: ConvertFromInts, RandomMirror, ToTensor, RandomVerticalFlip(0.3), RandomCrop, /255 --
    data_preprocess.py:13-46, transforms.py:356-362, 534-549, 738-748).  Returns the HR crop and its mask; the blur, the
    x4 downscale and the blur-kernel target are produced on the GPU for the whole batch by the trainer
    (csbsr_degrade with theta ~ U(0, 180) deg, sigma ~ U(0.2, 4) per sample, blur.py:128-179, 207-238)."""

    def __init__(self, cfg, image_dir=None, seg_dir=None, seed=None):
        from pathlib import Path
        self.image_dir = image_dir or cfg.DATASET.TRAIN_IMAGE_DIR
        self.seg_dir = seg_dir or cfg.DATASET.TRAIN_MASK_DIR
        self.fnames = sorted(path.name for path in Path(self.image_dir).glob("*.jpg"))
        self.size = tuple(cfg.INPUT.IMAGE_SIZE)
        self.vflip_p = 0.3
        for name, arg in cfg.DATASET.DATA_AUGMENTATION:
            if name == "RandomVerticalFlip":
                self.vflip_p = float(arg)
        self.rng = np.random.default_rng(seed)

    def __len__(self):
        return len(self.fnames)

    def __getitem__(self, i):
        from PIL import Image
        fname = self.fnames[i]
        img = np.array(Image.open(os.path.join(self.image_dir, fname))).astype(np.float32)            # H x W x 3
        seg = np.array(Image.open(os.path.join(self.seg_dir, fname))).astype(np.float32)
        if seg.ndim == 2:
            seg = seg[:, :, None]
        if self.rng.integers(2):                                       # RandomMirror
            img, seg = img[:, ::-1], seg[:, ::-1]
        if self.vflip_p <= self.rng.random():                          # RandomVerticalFlip: flips when p <= rand (sic)
            img, seg = img[::-1], seg[::-1]
        h, w = img.shape[:2]
        th, tw = self.size
        if h < th or w < tw:
            raise ValueError("image %s (%dx%d) is smaller than the crop %dx%d" % (fname, h, w, th, tw))
        y0 = int(self.rng.integers(0, h - th + 1))                     # RandomCrop.get_params
        x0 = int(self.rng.integers(0, w - tw + 1))
        img = np.ascontiguousarray(img[y0:y0 + th, x0:x0 + tw])
        seg = np.ascontiguousarray(seg[y0:y0 + th, x0:x0 + tw])
        return (torch.from_numpy(img).permute(2, 0, 1) / 255, torch.from_numpy(seg[:, :, :1]).permute(2, 0, 1) / 255)


class SyntheticCrackTestSet(torch.utils.data.Dataset):
    """Seeded synthetic crack images degraded on the fly by the device kernels (utils/synth.py + data/degrade.py)."""

    def __init__(self, n, size=448, start=0, seed=5):
        from ..utils import synth
        self.n, self.size, self.start = n, size, start
        self.params = synth.degradation_params(n, seed=seed)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        from ..utils import synth
        from . import degrade as G
        hr, mask = synth.crack_image(self.start + i, self.size)
        lr, kern = G.degrade(hr.unsqueeze(0), self.params[i:i + 1])
        return lr[0].cpu(), hr, mask, kern.cpu(), "synthetic_%05d.jpg" % (self.start + i)


def collate(items):
    lr, hr, mask, kern, names = zip(*items)
    return torch.stack(lr), torch.stack(hr), torch.stack(mask), torch.stack(kern), list(names)
