"""Shim for skimage.segmentation.find_boundaries (mode='inner', connectivity=1) -- parity unpinned
at this boundary: scikit-image is absent here, this restatement follows its documented algorithm."""
import numpy as np
from scipy import ndimage as ndi


def find_boundaries(label_img, connectivity=1, mode='thick', background=0):
    if label_img.dtype == bool:
        label_img = label_img.astype(np.uint8)
    fp = ndi.generate_binary_structure(label_img.ndim, connectivity)
    b = ndi.grey_dilation(label_img, footprint=fp) != ndi.grey_erosion(label_img, footprint=fp)
    if mode == 'inner':
        return b & (label_img != background)
    return b
