"""TEST INFRASTRUCTURE ONLY -- numpy oracle for the image path of the training augmentation
(reference dataset.py:48-105, ``PanoCorBonDataset.__getitem__``), "next" row f3 of SURVEY.md section 8.

Restates, with the reference line each step follows:
  * uint8 HWC image -> float32 in [0,1]                          dataset.py:53
  * stretch (misc/panostretch.pano_stretch, oracle/panostretch_ref.py)   dataset.py:69-82
  * horizontal flip                                               dataset.py:88-91
  * horizontal roll by dx pixels                                  dataset.py:94-98
  * gamma: img ** p in float32                                    dataset.py:101-105
  * HWC -> CHW float tensor                                       dataset.py:124
and the matching corner bookkeeping (x only; the 1-D boundary/corner label maths stays on the CPU in the product
as well).  The random draws of the reference (np.random.uniform / randint, dataset.py:71-81, 88, 95, 102-104) are
INPUTS here: (kx, ky, flip, dx, p).

Pinning: tests/golden/make_golden.py runs the REAL ``dataset.PanoCorBonDataset`` (only ``shapely`` is stubbed: it is
not installed and is used for occlusion labels only) on a synthetic image with a seeded np.random, replays the same
draws to recover (kx, ky, flip, dx, p) and stores the reference's tensor; tests/test_oracle.py asserts this
restatement reproduces it.
"""
import numpy as np

from . import panostretch_ref


def augment_image(img_u8, kx=None, ky=None, flip=False, dx=0, p=None, use_scipy=False):
    """img_u8: [H, W, 3] uint8.  Returns x [3, H, W] float32 exactly as dataset.py:124 builds it.
    kx/ky None = no stretch (self.stretch False); p None = no gamma."""
    img = np.array(img_u8, np.float32)[..., :3] / 255.                               # dataset.py:53
    if kx is not None:
        img, _ = panostretch_ref.pano_stretch(img, np.zeros((1, 2), np.float32), kx, ky, use_scipy=use_scipy)   # :82
    if flip:
        img = np.flip(img, axis=1)                                                    # :89
    if dx:
        img = np.roll(img, dx, axis=1)                                                # :96
    if p is not None:
        img = img ** p                                                                # :105 (float32 ** python float)
    return np.ascontiguousarray(img.transpose([2, 0, 1])).astype(np.float32)          # :124


def augment_corners(cor, H, W, kx=None, ky=None, flip=False, dx=0):
    """cor: [N, 2] (x, y) float32 corner list; the x/y bookkeeping of dataset.py:82, 91, 98."""
    cor = np.array(cor, dtype=np.float32, copy=True)
    if kx is not None:
        cor = panostretch_ref.stretch_corners(cor, H, W, kx, ky)                      # :82 (returns float64)
    if flip:
        cor[:, 0] = W - 1 - cor[:, 0]                                                 # :91
    if dx:
        cor[:, 0] = (cor[:, 0] + dx) % W                                              # :98
    return cor
