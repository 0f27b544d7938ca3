"""Device-side training augmentation (reference CrackDataSet.__getitem__, model/data/crack_dataset.py:42-48, TrainTransforms of
config_csbsr_pspnet.yaml via data_preprocess.py:13-46): the JPEGs are decoded on the host (PIL), the uint8 HWC arrays are
uploaded once, and ONE kernel (csbsr_crop_flip_u8) does ConvertFromInts + RandomMirror + RandomVerticalFlip + RandomCrop +
ToTensor + /255 for the whole batch.  The random draws stay on the host in the reference's order (mirror, vertical flip, crop)."""
import numpy as np
import torch

from .. import _lib


def draw_params(shapes, size, rng, vflip_p=0.3):
    """Per image (H, W): (y0, x0, hflip, vflip) -- RandomMirror (transforms.py:356-362), RandomVerticalFlip flips when
    p <= rand (sic, transforms.py:738-748), RandomCrop.get_params (transforms.py:534-549)."""
    th, tw = size
    out = np.zeros((len(shapes), 4), dtype=np.int32)
    for i, (h, w) in enumerate(shapes):
        if h < th or w < tw:
            raise ValueError("image %d (%dx%d) is smaller than the crop %dx%d" % (i, h, w, th, tw))
        out[i, 2] = int(rng.integers(2))
        out[i, 3] = int(vflip_p <= rng.random())
        out[i, 0] = int(rng.integers(0, h - th + 1))
        out[i, 1] = int(rng.integers(0, w - tw + 1))
    return out


def crop_flip_batch(images, params, size, c_out=None, device=None):
    """images: list of uint8 numpy / torch arrays [H, W, C] (or [H, W]); params int32 [B, 4] -> fp32 [B, c_out, th, tw] in [0, 1]."""
    dev = device or torch.device("cuda", torch.cuda.current_device())
    th, tw = size
    ts = []
    for im in images:
        t = torch.as_tensor(np.ascontiguousarray(im) if isinstance(im, np.ndarray) else im)
        if t.dim() == 2:
            t = t.unsqueeze(2)
        assert t.dtype == torch.uint8
        ts.append(t.contiguous().to(dev, non_blocking=True))
    b = len(ts)
    c_out = c_out or ts[0].shape[2]
    ptrs = torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64).to(dev)
    dims = torch.tensor([list(t.shape) for t in ts], dtype=torch.int32).to(dev)
    prm = torch.as_tensor(np.asarray(params, dtype=np.int32)).to(dev)
    out = torch.empty((b, c_out, th, tw), dtype=torch.float32, device=dev)
    import ctypes as C
    _lib.check(_lib.lib().csbsr_crop_flip_u8(ptrs.data_ptr(), dims.data_ptr(), prm.data_ptr(), out.data_ptr(), b, c_out, th, tw,
                                             C.c_float(255.0), _lib.stream_ptr()), "csbsr_crop_flip_u8")
    _lib.count_launch("csbsr_crop_flip_u8")
    return out
