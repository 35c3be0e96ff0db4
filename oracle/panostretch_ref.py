"""TEST INFRASTRUCTURE ONLY -- numpy oracle for ``misc/panostretch.pano_stretch``.

Restates reference misc/panostretch.py:81-117 (with uv_meshgrid :6-11, _uv_tri :14-20,
coorx2u/coory2v/u2coorx/v2coory :28-41) and the one third-party routine it calls,
``scipy.ndimage.map_coordinates(order<=1, mode='wrap')`` (call site panostretch.py:99-102;
reference pins scipy==1.3.0 in environment.yml:8, the only testable version is scipy 1.18.1).
SciPy's published algorithm for that call (ndimage/src/ni_interpolation.c, ``map_coordinate`` with
NI_EXTEND_WRAP, then linear interpolation in double precision, result cast to the input dtype):

  * the *legacy* 'wrap' maps a coordinate into [0, n-1] with period n-1 (not n):
        c < 0   ->  c + (n-1) * (floor(-c / (n-1)) + 1)
        c > n-1 ->  c - (n-1) * floor(c / (n-1))
  * order 1: i0 = floor(c), t = c - i0, value = (1-t)*a[i0] + t*a[i0+1] per axis (separable);
    an index that falls outside [0, n-1] (only i0+1 == n, whose weight is exactly 0) is
    folded back with the same period, so it never contributes.
  * order 0: index floor(c + 0.5).

Pinning: no reference test exists; tests/golden/make_golden.py runs the real
``misc.panostretch.pano_stretch`` (scipy 1.18.1) in the build container and stores outputs in
tests/golden/panostretch_*.npz; tests/test_oracle.py asserts this restatement reproduces them.
"""
import numpy as np


def _legacy_wrap(c, n):
    c = np.array(c, dtype=np.float64, copy=True)
    if n <= 1:
        return np.zeros_like(c)
    sz = float(n - 1)
    lo = c < 0
    c[lo] = c[lo] + sz * (np.floor(-c[lo] / sz) + 1.0)
    hi = c > sz
    c[hi] = c[hi] - sz * np.floor(c[hi] / sz)
    return c


def stretch_coords(h, w, kx, ky):
    """panostretch.py:91-96: per-pixel source coordinates (refy, refx), fp64, pixel units."""
    xs = ((np.arange(w, dtype=np.float64) + 0.5) / w - 0.5) * 2 * np.pi       # :9  u
    ys = ((np.arange(h, dtype=np.float64) + 0.5) / h - 0.5) * np.pi           # :10 v
    sin_u, cos_u = np.sin(xs)[None, :], np.cos(xs)[None, :]                    # :17-18
    tan_v = np.tan(ys)[:, None]                                                # :19
    u0 = np.arctan2(sin_u * kx / ky, cos_u)                                    # :92
    v0 = np.arctan(tan_v * np.sin(u0) / sin_u * ky)                            # :93
    refx = (u0 / (2 * np.pi) + 0.5) * w - 0.5                                  # :95
    refy = (v0 / np.pi + 0.5) * h - 0.5                                        # :96
    return refy, np.broadcast_to(refx, refy.shape)


def map_coordinates_wrap(a, refy, refx, order=1):
    """scipy.ndimage.map_coordinates(a, [refy, refx], order=order, mode='wrap') for 2-D ``a``."""
    h, w = a.shape
    cy = _legacy_wrap(refy, h)
    cx = _legacy_wrap(refx, w)
    src = a.astype(np.float64)
    if order == 0:
        iy = np.floor(cy + 0.5).astype(np.int64)
        ix = np.floor(cx + 0.5).astype(np.int64)
        return src[iy, ix].astype(a.dtype)
    if order != 1:
        raise NotImplementedError('the hot path uses order 0/1 only')
    y0 = np.floor(cy).astype(np.int64)
    x0 = np.floor(cx).astype(np.int64)
    ty = cy - y0
    tx = cx - x0
    y1 = np.where(y0 + 1 > h - 1, (y0 + 1) - (h - 1), y0 + 1) if h > 1 else y0
    x1 = np.where(x0 + 1 > w - 1, (x0 + 1) - (w - 1), x0 + 1) if w > 1 else x0
    # scipy accumulates coefficient * value over the 2x2 support in double precision
    val = ((1 - ty) * (1 - tx)) * src[y0, x0] + ((1 - ty) * tx) * src[y0, x1] \
        + (ty * (1 - tx)) * src[y1, x0] + (ty * tx) * src[y1, x1]
    return val.astype(a.dtype)


def stretch_corners(corners, h, w, kx, ky):
    """panostretch.py:104-115: closed-form transform of the [N,2] (x,y) corner list, fp64."""
    corners = np.asarray(corners)
    u0 = ((corners[:, 0] + 0.5) / w - 0.5) * 2 * np.pi                         # :28-29
    v0 = ((corners[:, 1] + 0.5) / h - 0.5) * np.pi                             # :32-33
    u = np.arctan2(np.sin(u0) * ky / kx, np.cos(u0))                           # :107
    c2 = (np.sin(u0) * ky) ** 2 + (np.cos(u0) * kx) ** 2                       # :108
    v = np.arctan2(np.sin(v0), np.cos(v0) * np.sqrt(c2))                       # :109-111
    return np.stack([(u / (2 * np.pi) + 0.5) * w - 0.5, (v / np.pi + 0.5) * h - 0.5], axis=-1)


def pano_stretch(img, corners, kx, ky, order=1, use_scipy=False):
    """Same signature and return convention as reference panostretch.py:81.  ``use_scipy=True`` samples with the real
    ``scipy.ndimage.map_coordinates`` (the third-party routine the reference calls at :99-102) instead of the numpy
    restatement above: that is the reference's actual CPU cost, used by bench.py's CPU baseline."""
    h, w = img.shape[:2]
    refy, refx = stretch_coords(h, w, kx, ky)
    if use_scipy:
        from scipy.ndimage import map_coordinates
        out = np.stack([map_coordinates(img[..., i], [refy, refx], order=order, mode='wrap')
                        for i in range(img.shape[-1])], axis=-1)
        return out, stretch_corners(corners, h, w, kx, ky)
    out = np.stack([map_coordinates_wrap(img[..., i], refy, refx, order)
                    for i in range(img.shape[-1])], axis=-1)
    return out, stretch_corners(corners, h, w, kx, ky)
