"""Drop-in ``rotatePanorama`` (reference misc/pano_lsd_align.py:125-171) on the CUDA library ("next" row f4).

Same signature and return value as the reference: ``rotatePanorama(img[H,W,C], vp=None, R=None) -> float64 [H,W,C]``
(preprocess.py:65-66 calls it with the vanishing points, ``vp[2::-1]``).  The 3x3 inverse is taken on the host in fp64
(the reference solves ``R x = xyz`` per pixel, :146; multiplying by the inverse differs by ~1e-16); everything per pixel
-- angles, rotation, asin, the implicit one-pixel padding of :156-168 and the bilinear sample -- runs in one kernel.
The rest of the file (LSD line detection, vanishing-point voting) needs ``pylsd`` and is out of scope.
"""
import ctypes

import numpy as np

from .. import _lib


def _rinv(vp, R):
    if R is None:
        if vp is None:
            raise ValueError('rotatePanorama needs vp or R')
        R = np.linalg.inv(np.asarray(vp, np.float64).T)                 # pano_lsd_align.py:143
    rinv = np.ascontiguousarray(np.linalg.inv(np.asarray(R, np.float64)))
    return (ctypes.c_double * 9)(*rinv.reshape(-1))


def rotate_panorama_batch(imgs, vp=None, R=None, out=None):
    """Device-resident batch: imgs is a CUDA float32 / float64 tensor [N, H, W, C]; returns CUDA float64 [N, H, W, C]."""
    import torch
    if not (isinstance(imgs, torch.Tensor) and imgs.is_cuda and imgs.dim() == 4 and
            imgs.dtype in (torch.float32, torch.float64)):
        raise TypeError('rotate_panorama_batch expects a CUDA float32/float64 tensor [N, H, W, C]')
    imgs = imgs.contiguous()
    n, h, w, c = imgs.shape
    if out is None:
        out = torch.empty(n, h, w, c, device=imgs.device, dtype=torch.float64)
    stream = torch.cuda.current_stream(imgs.device).cuda_stream
    with torch.cuda.device(imgs.device):
        _lib.check(_lib.lib().hn_rotate_panorama(imgs.data_ptr(), 1 if imgs.dtype == torch.float64 else 0, out.data_ptr(),
                                                 n, h, w, c, _rinv(vp, R), stream), 'hn_rotate_panorama')
    return out


def rotatePanorama(img, vp=None, R=None):
    """img: [H, W, C] numpy array (float32 / float64; other dtypes are converted to float64 like the reference's
    ``imgNew[...] = img`` assignment does).  Returns float64 [H, W, C]."""
    import torch
    img = np.asarray(img)
    if img.ndim != 3:
        raise ValueError('img must be [H, W, C]')
    if img.dtype not in (np.float32, np.float64):
        img = img.astype(np.float64)
    if not torch.cuda.is_available():
        raise RuntimeError('horizonnet_b200 has no CPU path: rotatePanorama needs a B200 (cuda) device')
    d = torch.from_numpy(np.ascontiguousarray(img)).cuda().unsqueeze(0)
    return rotate_panorama_batch(d, vp, R)[0].cpu().numpy()
