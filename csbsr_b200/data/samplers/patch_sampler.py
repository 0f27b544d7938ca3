"""SplitPatch / JointPatch with the reference's call convention (model/data/samplers/patch_sampler.py:15-50), running on the
device through csbsr_patch_split / csbsr_patch_join (csrc/glue.cu).  test.py uses them for test images larger than
INPUT.IMAGE_SIZE: every image is cut into IMAGE_SIZE / scale LR patches, the networks run per patch, and the SR / segmentation
outputs are joined back before the metrics (model/engine/inference.py:80-91)."""
import numpy as np
import torch

from ... import _lib


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


class SplitPatch:
    def __init__(self, batch_size, ch, patch_sizeh, patch_sizew):
        self.kc, self.kh, self.kw = ch, patch_sizeh, patch_sizew
        self.batch_size = batch_size

    def __call__(self, x):
        """x fp32 [C, H, W] -> (patches [ny * nx, C, kh, kw], unfold_shape = [batch_size, 1, ny, nx, C, kh, kw])."""
        x = x.to(device=_dev(), dtype=torch.float32).contiguous()
        c, h, w = x.shape
        assert c == self.kc, "SplitPatch: channel count differs from the patch channel count"
        ny, nx = h // self.kh, w // self.kw
        out = torch.empty((ny * nx, c, self.kh, self.kw), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().csbsr_patch_split(x.data_ptr(), out.data_ptr(), 1, c, h, w, self.kh, self.kw, _lib.stream_ptr()),
                   "csbsr_patch_split")
        _lib.count_launch("csbsr_patch_split")
        return out, np.array([self.batch_size, 1, ny, nx, c, self.kh, self.kw])


class JointPatch:
    def __call__(self, patches, unfold_shape, batch_size=-1):
        """patches [B * ny * nx, C, ph, pw] + unfold_shape [_, 1, ny, nx, C, ph, pw] -> [B, C, ny * ph, nx * pw]."""
        patches = patches.to(device=_dev(), dtype=torch.float32).contiguous()
        _, sc, ny, nx, c, ph, pw = [int(v) for v in unfold_shape]
        assert sc == 1 and patches.shape[1:] == (c, ph, pw) and patches.shape[0] % (ny * nx) == 0
        b = patches.shape[0] // (ny * nx)
        out = torch.empty((b, c, ny * ph, nx * pw), dtype=torch.float32, device=patches.device)
        _lib.check(_lib.lib().csbsr_patch_join(patches.data_ptr(), out.data_ptr(), b, c, ny * ph, nx * pw, ph, pw, _lib.stream_ptr()),
                   "csbsr_patch_join")
        _lib.count_launch("csbsr_patch_join")
        return out
