"""On-device training augmentation, image path ("next" row f3; reference dataset.py:48-134).

The reference's ``PanoCorBonDataset.__getitem__`` loads a uint8 panorama, converts it to float32, and -- per sample, in
a DataLoader worker process -- stretches it (``pano_stretch``, ~70 ms on one core), flips it, rolls it, applies a
gamma curve and transposes it to CHW, materialising a full float32 image after every step (dataset.py:53, 69-105,
124).  ``augment_batch`` does the image part of all of that in ONE gather kernel per batch on the GPU, from the uint8
upload (1.5 MB per panorama instead of 6.3 MB) straight to the float32 ``[N, 3, H, W]`` network input.

What stays on the host, exactly as in the reference (small, label-side, CPU glue): the random draws
(``draw_params``, the same ``np.random`` calls in the same order as dataset.py:71-81, 88, 95, 102-104) and the corner
bookkeeping (``augment_corners``: dataset.py:82, 91, 98); the 1-D boundary / corner targets (``cor_2_1d``, ``cdist``)
are out of scope.  No CPU fallback: without the library / a GPU the call raises.
"""
import ctypes

import numpy as np

from . import _lib
from .misc.panostretch import _stretch_corners


def _uv2xy(u, v, z=-50):                                   # reference misc/panostretch.py:44-48
    c = z / np.tan(v)
    return c * np.cos(u), c * np.sin(u)


def cor2xybound(cor, w=1024, h=512):
    """Stretch-factor bounds of a room (reference dataset.py:198-217; coorx2u / coory2v use their default 1024 x 512
    there as well)."""
    corU, corB = cor[0::2], cor[1::2]
    zU = -50
    u = ((corU[:, 0] + 0.5) / w - 0.5) * 2 * np.pi         # panostretch.py:28-29
    vU = ((corU[:, 1] + 0.5) / h - 0.5) * np.pi            # :32-33
    vB = ((corB[:, 1] + 0.5) / h - 0.5) * np.pi
    x, y = _uv2xy(u, vU, z=zU)
    c = np.sqrt(x ** 2 + y ** 2)
    zB = c * np.tan(vB)
    xmin, xmax = x.min(), x.max()
    ymin, ymax = y.min(), y.max()
    S = 3 / abs(zB.mean() - zU)
    dx = [abs(xmin * S), abs(xmax * S)]
    dy = [abs(ymin * S), abs(ymax * S)]
    return min(dx), min(dy), max(dx), max(dy)


def draw_params(cor, W, stretch=True, flip=True, rotate=True, gamma=True, max_stretch=2.0, rng=np.random):
    """The random draws of dataset.py:69-105 in the reference's order -> dict(kx, ky, flip, dx, p); kx = ky = None
    without stretch, p = None without gamma.  ``rng`` = np.random (module) or a RandomState."""
    out = dict(kx=None, ky=None, flip=False, dx=0, p=None)
    if stretch:
        xmin, ymin, xmax, ymax = cor2xybound(cor)                       # dataset.py:70
        kx = rng.uniform(1.0, max_stretch)                              # :71
        ky = rng.uniform(1.0, max_stretch)                              # :72
        if rng.randint(2) == 0:                                         # :73-76
            kx = max(1 / kx, min(0.5 / xmin, 1.0))
        else:
            kx = min(kx, max(10.0 / xmax, 1.0))
        if rng.randint(2) == 0:                                         # :77-80
            ky = max(1 / ky, min(0.5 / ymin, 1.0))
        else:
            ky = min(ky, max(10.0 / ymax, 1.0))
        out['kx'], out['ky'] = float(kx), float(ky)
    if flip and rng.randint(2) == 0:                                    # :88
        out['flip'] = True
    if rotate:
        out['dx'] = int(rng.randint(W))                                 # :95
    if gamma:
        p = rng.uniform(1, 2)                                           # :102
        if rng.randint(2) == 0:                                         # :103-104
            p = 1 / p
        out['p'] = float(p)
    return out


def augment_corners(cor, H, W, kx=None, ky=None, flip=False, dx=0):
    """Corner list bookkeeping of dataset.py:82, 91, 98 (host, numpy)."""
    cor = np.array(cor, dtype=np.float32, copy=True)
    if kx is not None:
        cor = _stretch_corners(cor, H, W, kx, ky)
    if flip:
        cor[:, 0] = W - 1 - cor[:, 0]
    if dx:
        cor[:, 0] = (cor[:, 0] + dx) % W
    return cor


def augment_batch(imgs_u8, kx=None, ky=None, flip=None, dx=None, gamma=None, out=None, device=None):
    """imgs_u8: [N, H, W, 3] uint8 -- a CUDA tensor, or a numpy array / CPU tensor (uploaded as uint8).
    kx, ky, flip, dx, gamma: length-N sequences (None entries / None = that augmentation off for the image / batch).
    Returns x: CUDA float32 [N, 3, H, W], the tensor dataset.py:124 builds per sample."""
    import torch
    if not isinstance(imgs_u8, torch.Tensor):
        imgs_u8 = torch.from_numpy(np.ascontiguousarray(imgs_u8))
    if imgs_u8.dtype != torch.uint8 or imgs_u8.dim() != 4 or imgs_u8.shape[3] != 3:
        raise TypeError('augment_batch expects uint8 images [N, H, W, 3] (dataset.py:53 reads RGB uint8)')
    if not imgs_u8.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError('horizonnet_b200 has no CPU path: augment_batch needs a B200 (cuda) device')
        imgs_u8 = imgs_u8.to(device if device is not None else 'cuda', non_blocking=True)
    imgs_u8 = imgs_u8.contiguous()
    n, h, w, _ = imgs_u8.shape
    if out is None:
        out = torch.empty(n, 3, h, w, device=imgs_u8.device, dtype=torch.float32)

    def arr(ctype, seq, none_value):
        if seq is None:
            return None
        vals = [none_value if v is None else v for v in seq]
        if len(vals) != n:
            raise ValueError('per-image parameter lists must have N entries')
        return (ctype * n)(*vals)
    kxa = arr(ctypes.c_double, kx, -1.0)
    kya = arr(ctypes.c_double, ky, -1.0)
    fla = arr(ctypes.c_int, None if flip is None else [1 if f else 0 for f in flip], 0)
    dxa = arr(ctypes.c_int, dx, 0)
    gaa = arr(ctypes.c_float, gamma, -1.0)
    stream = torch.cuda.current_stream(imgs_u8.device).cuda_stream
    with torch.cuda.device(imgs_u8.device):
        _lib.check(_lib.lib().hn_augment(imgs_u8.data_ptr(), out.data_ptr(), n, h, w, kxa, kya, fla, dxa, gaa, stream),
                   'hn_augment')
    return out


def bench_aux(dev, peaks, n_img=64, reps=10):
    """bench.py aux leg: the fused augmentation pass on 64 uint8 panoramas with per-image random parameters, against
    the HBM roofline (algorithmic bytes: H*W*3 in + H*W*3*4 out per panorama)."""
    import torch
    rs = np.random.RandomState(5)
    imgs = torch.randint(0, 256, (n_img, 512, 1024, 3), dtype=torch.uint8, device=dev)
    kx = [float(v) for v in rs.uniform(0.5, 2.0, n_img)]
    ky = [float(v) for v in rs.uniform(0.5, 2.0, n_img)]
    flip = [bool(v) for v in rs.randint(0, 2, n_img)]
    dx = [int(v) for v in rs.randint(0, 1024, n_img)]
    gam = [float(v) for v in rs.uniform(0.5, 2.0, n_img)]
    out = torch.empty(n_img, 3, 512, 1024, device=dev)
    for _ in range(3):
        augment_batch(imgs, kx, ky, flip, dx, gam, out=out)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(reps):
        augment_batch(imgs, kx, ky, flip, dx, gam, out=out)
    a1.record()
    torch.cuda.synchronize()
    ms = a0.elapsed_time(a1) / reps
    bytes_per = 512 * 1024 * 3 * 5
    gbs = n_img * bytes_per / (ms * 1e-3) / 1e9
    res = {'augment_fused': {'panos_per_s': round(n_img / (ms * 1e-3), 1), 'achieved_gbs': round(gbs, 1),
                             'peak_gbs': peaks['hbm_gbs'], 'frac': round(gbs / peaks['hbm_gbs'], 4), 'bytes_per_pano': bytes_per,
                             'what': 'uint8 HWC -> stretch + flip + roll + gamma -> float32 CHW in one pass (dataset.py:53,69-105,124), 64 panos per launch'}}
    try:
        from .misc.pano_lsd_align import rotate_panorama_batch
        img = torch.rand(16, 512, 1024, 3, device=dev)
        q, _ = np.linalg.qr(np.random.RandomState(3).randn(3, 3))
        o = None
        for _ in range(2):
            o = rotate_panorama_batch(img, R=q, out=o)
        a0.record()
        for _ in range(reps):
            rotate_panorama_batch(img, R=q, out=o)
        a1.record()
        torch.cuda.synchronize()
        ms = a0.elapsed_time(a1) / reps
        bytes_per = 512 * 1024 * 3 * 12
        gbs = 16 * bytes_per / (ms * 1e-3) / 1e9
        res['rotate_panorama'] = {'panos_per_s': round(16 / (ms * 1e-3), 1), 'achieved_gbs': round(gbs, 1),
                                  'frac': round(gbs / peaks['hbm_gbs'], 4), 'bytes_per_pano': bytes_per,
                                  'what': 'rotatePanorama (pano_lsd_align.py:125-171), float32 in / float64 out, 16 panos per launch'}
    except Exception as e:
        res['rotate_panorama'] = {'error': str(e)}
    return res
