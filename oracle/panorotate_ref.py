"""TEST INFRASTRUCTURE ONLY -- numpy oracle for ``misc/pano_lsd_align.rotatePanorama`` ("next" row f4).

Restates reference misc/pano_lsd_align.py:125-171 (rotatePanorama) with the helpers it calls:
uv2xyzN :72-80 (planeID 1), xyz2uvN :53-69 (planeID 1), warpImageFast :101-122, whose sampler is
``scipy.ndimage.map_coordinates(order=1)`` with the default mode='constant', cval=0 (bilinear in double precision,
taps outside the array contribute 0; call site :116-119).  As in oracle/panostretch_ref.py the scipy routine is
restated in numpy (``use_scipy=True`` calls the real one).

Pinning: tests/golden/make_golden.py runs the REAL rotatePanorama and stores its output; tests/test_oracle.py asserts
this restatement reproduces it (<= 1e-12).
"""
import numpy as np


def rotation_coords(H, W, R):
    """pano_lsd_align.py:133-154: source pixel coordinates (1-based, Px along W, Py along H) of every target pixel."""
    TX, TY = np.meshgrid(range(1, W + 1), range(1, H + 1))                           # :133
    TX = TX.reshape(-1, 1, order='F')
    TY = TY.reshape(-1, 1, order='F')
    ANGx = (TX - W / 2 - 0.5) / W * np.pi * 2                                        # :136
    ANGy = -(TY - H / 2 - 0.5) / H * np.pi                                           # :137
    xyz = np.zeros((ANGx.shape[0], 3))
    xyz[:, 0] = (np.cos(ANGy) * np.sin(ANGx))[:, 0]                                  # uv2xyzN planeID=1, :76-79
    xyz[:, 1] = (np.cos(ANGy) * np.cos(ANGx))[:, 0]
    xyz[:, 2] = np.sin(ANGy)[:, 0]
    old = np.linalg.solve(R, xyz.T).T                                                # :146
    normXY = np.sqrt(old[:, [0]] ** 2 + old[:, [1]] ** 2)                            # xyz2uvN :57-68
    normXY[normXY < 0.000001] = 0.000001
    normXYZ = np.sqrt(old[:, [0]] ** 2 + old[:, [1]] ** 2 + old[:, [2]] ** 2)
    v = np.arcsin(old[:, [2]] / normXYZ)
    u = np.arcsin(old[:, [0]] / normXY)
    valid = (old[:, [1]] < 0) & (u >= 0)
    u[valid] = np.pi - u[valid]
    valid = (old[:, [1]] < 0) & (u <= 0)
    u[valid] = -np.pi - u[valid]
    u[np.isnan(u)] = 0
    Px = (u[:, 0] + np.pi) / (2 * np.pi) * W + 0.5                                   # :149
    Py = (-v[:, 0] + np.pi / 2) / np.pi * H + 0.5                                    # :150
    return Px.reshape(H, W, order='F'), Py.reshape(H, W, order='F')


def padded_image(img):
    """pano_lsd_align.py:156-168, including the last-row quirk at :163 (its right half copies img[0], not img[-1])."""
    H, W, C = img.shape
    p = np.zeros((H + 2, W + 2, C), np.float64)
    p[1:-1, 1:-1, :] = img
    p[1:-1, 0, :] = img[:, -1, :]
    p[1:-1, -1, :] = img[:, 0, :]
    p[0, 1:W // 2 + 1, :] = img[0, W - 1:W // 2 - 1:-1, :]
    p[0, W // 2 + 1:-1, :] = img[0, W // 2 - 1::-1, :]
    p[-1, 1:W // 2 + 1, :] = img[-1, W - 1:W // 2 - 1:-1, :]
    p[-1, W // 2 + 1:-1, :] = img[0, W // 2 - 1::-1, :]
    p[0, 0, :] = img[0, 0, :]
    p[-1, -1, :] = img[-1, -1, :]
    p[0, -1, :] = img[0, -1, :]
    p[-1, 0, :] = img[-1, 0, :]
    return p


def _bilinear_constant(a, cy, cx):
    """scipy map_coordinates(a, [cy, cx], order=1) with mode='constant', cval=0 for a 2-D double array."""
    h, w = a.shape
    y0 = np.floor(cy).astype(np.int64)
    x0 = np.floor(cx).astype(np.int64)
    ty = cy - y0
    tx = cx - x0
    out = np.zeros(cy.shape, np.float64)
    for dy, wy in ((0, 1 - ty), (1, ty)):
        for dxx, wx in ((0, 1 - tx), (1, tx)):
            yy, xx = y0 + dy, x0 + dxx
            ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
            val = np.where(ok, a[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)], 0.0)
            out = out + (wy * wx) * val
    # scipy returns cval for coordinates outside [0, n-1] altogether
    inside = (cy >= 0) & (cy <= h - 1) & (cx >= 0) & (cx <= w - 1)
    return np.where(inside, out, 0.0)


def warp_image_fast(im, XX, YY, use_scipy=False):
    """pano_lsd_align.py:101-122."""
    minX = max(1., np.floor(XX.min()) - 1)
    minY = max(1., np.floor(YY.min()) - 1)
    maxX = min(im.shape[1], np.ceil(XX.max()) + 1)
    maxY = min(im.shape[0], np.ceil(YY.max()) + 1)
    im = im[int(round(minY - 1)):int(round(maxY)), int(round(minX - 1)):int(round(maxX))]
    cy, cx = YY - minY, XX - minX
    if use_scipy:
        from scipy.ndimage import map_coordinates
        return np.stack([map_coordinates(im[..., c], [cy.reshape(-1), cx.reshape(-1)], order=1).reshape(XX.shape)
                         for c in range(im.shape[-1])], axis=-1)
    return np.stack([_bilinear_constant(im[..., c], cy, cx) for c in range(im.shape[-1])], axis=-1)


def rotate_panorama(img, vp=None, R=None, use_scipy=False):
    """Same signature and return value as reference rotatePanorama (float64 [H, W, C])."""
    H, W, C = img.shape
    if R is None:
        R = np.linalg.inv(vp.T)                                                      # :143
    Px, Py = rotation_coords(H, W, R)
    return warp_image_fast(padded_image(img), Px + 1, Py + 1, use_scipy)             # :170
