"""Drop-in ``pano_stretch`` (reference misc/panostretch.py:81-117) on the CUDA library.

Same signature and return convention as the reference:
    pano_stretch(img[H,W,C] float32|float64, corners[N,2], kx, ky, order=1) -> (img'[H,W,C] same dtype, corners'[N,2] f64)
The image warp (the hot loop: arctan per pixel + scipy map_coordinates per channel, ~67 ms/img on
one CPU core) runs in one fused gather kernel; the corner transform is a closed form on N <= a few
dozen points and stays on the host in fp64 exactly as the reference computes it (:104-115).
``pano_stretch_batch`` is the device-resident form for augmentation loops (no host round trip).
No CPU fallback: without the library / a GPU the call raises.
"""
import ctypes

import numpy as np

from .. import _lib


def _stretch_corners(corners, h, w, kx, ky):
    corners = np.asarray(corners)
    u0 = ((corners[:, 0] + 0.5) / w - 0.5) * 2 * np.pi          # coorx2u, panostretch.py:28-29
    v0 = ((corners[:, 1] + 0.5) / h - 0.5) * np.pi              # coory2v, :32-33
    u = np.arctan2(np.sin(u0) * ky / kx, np.cos(u0))            # :107
    c2 = (np.sin(u0) * ky) ** 2 + (np.cos(u0) * kx) ** 2        # :108
    v = np.arctan2(np.sin(v0), np.cos(v0) * np.sqrt(c2))        # :109-111
    x = (u / (2 * np.pi) + 0.5) * w - 0.5                       # u2coorx, :36-37
    y = (v / np.pi + 0.5) * h - 0.5                             # v2coory, :40-41
    return np.stack([x, y], axis=-1)


def pano_stretch(img, corners, kx, ky, order=1):
    """img: [H, W, C] float32 numpy array; corners: [N, 2] (x, y); kx/ky: stretch along
    front-back / left-right; order: 0 nearest, 1 bilinear."""
    img = np.asarray(img)
    if img.ndim != 3:
        raise ValueError('img must be [H, W, C]')
    if img.dtype not in (np.float32, np.float64):
        raise TypeError('horizonnet_b200.pano_stretch handles float32 images (the training path, reference dataset.py:53) and '
                        'float64 images (the reference CLI, misc/panostretch.py:171); got ' + str(img.dtype))
    if order not in (0, 1):
        raise NotImplementedError('order 0 / 1 only (the reference callers use order=1)')
    h, w, c = img.shape
    src = np.ascontiguousarray(img)
    out = np.empty_like(src)
    kxa = (ctypes.c_double * 1)(float(kx))
    kya = (ctypes.c_double * 1)(float(ky))
    fn = _lib.lib().hn_pano_stretch_host_f64 if img.dtype == np.float64 else _lib.lib().hn_pano_stretch_host
    _lib.check(fn(src.ctypes.data, out.ctypes.data, 1, h, w, c, kxa, kya, int(order)), 'hn_pano_stretch_host')
    return out, _stretch_corners(corners, h, w, kx, ky)


def pano_stretch_batch(imgs, kx, ky, order=1, out=None):
    """Device-resident batch: imgs is a CUDA float32 tensor [N, H, W, C]; kx, ky are length-N
    sequences.  Returns a new CUDA tensor (or fills ``out``)."""
    import torch
    if not (isinstance(imgs, torch.Tensor) and imgs.is_cuda and imgs.dtype in (torch.float32, torch.float64) and imgs.dim() == 4):
        raise TypeError('pano_stretch_batch expects a CUDA float32 / float64 tensor [N, H, W, C]')
    imgs = imgs.contiguous()
    n, h, w, c = imgs.shape
    if out is None:
        out = torch.empty_like(imgs)
    kxa = (ctypes.c_double * n)(*[float(v) for v in kx])
    kya = (ctypes.c_double * n)(*[float(v) for v in ky])
    stream = torch.cuda.current_stream(imgs.device).cuda_stream
    with torch.cuda.device(imgs.device):
        fn = _lib.lib().hn_pano_stretch_f64 if imgs.dtype == torch.float64 else _lib.lib().hn_pano_stretch
        _lib.check(fn(imgs.data_ptr(), out.data_ptr(), n, h, w, c, kxa, kya, int(order), stream), 'hn_pano_stretch')
    return out
