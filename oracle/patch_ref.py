"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's patch split / join and training augmentation.

  split_patch / joint_patch   SplitPatch.__call__ / JointPatch.__call__   model/data/samplers/patch_sampler.py:15-50
  crop_flip                   RandomMirror / RandomVerticalFlip / RandomCrop / ToTensor / 255 as CrackDataSet applies them
                              (model/data/crack_dataset.py:42-48; transforms.py:356-362, 534-549, 738-748)
Pinned against the unmodified reference classes by tests/test_oracle_cpu.py::test_patch_oracle_matches_reference (build
container only: the reference is imported through oracle/ref_harness.py when present)."""
import numpy as np
import torch


def split_patch(x, batch_size, ch, ph, pw):
    patches = x.unfold(0, ch, ch).unfold(1, ph, ph).unfold(2, pw, pw)
    shape = patches.size()
    return patches.contiguous().view(-1, ch, ph, pw), np.append(batch_size, np.array(shape))


def joint_patch(patches, unfold_shape):
    u = list(unfold_shape)
    u[0] = -1
    p = patches.view(*u)
    c, h, w = u[1] * u[4], u[2] * u[5], u[3] * u[6]
    return p.permute(0, 1, 4, 2, 5, 3, 6).contiguous().view(-1, c, h, w)


def crop_flip(img_u8, y0, x0, hflip, vflip, th, tw):
    a = np.asarray(img_u8).astype(np.float32)
    if a.ndim == 2:
        a = a[:, :, None]
    if hflip:
        a = a[:, ::-1]
    if vflip:
        a = a[::-1]
    a = np.ascontiguousarray(a[y0:y0 + th, x0:x0 + tw])
    return torch.from_numpy(a).permute(2, 0, 1) / 255
