"""Device-side test-time augmentation around the forward (reference inference.py:32-62, 77-93).

``tta_forward(net, x, flip, rotate)`` replaces the first half of the reference's ``inference()``:
``augment`` (numpy flip / roll on the host), ``net(x.to(device))``, two ``.cpu()`` round trips, ``sigmoid``,
``augment_undo(...).mean(0)`` and the boundary -> pixel-row conversion with clipping.  Here one panorama is
uploaded once; the views, the forward and the merge all run on the device.  The CPU post-processing that
follows in the reference (post_proc.*, peak finding, polygon checks) is out of scope and takes these outputs
as they are.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._spec import PANO_H, PANO_W


def tta_forward(net, x, flip=False, rotate=()):
    """x: [1, 3, 512, 1024] float tensor (any device).  Returns (y_bon_ [2,1024] pixel rows, y_cor_ [1024]) as
    float32 numpy arrays -- the values the reference holds after inference.py:93 (before post-processing)."""
    if x.dim() != 4 or x.shape[0] != 1 or x.shape[1] != 3:
        raise ValueError('tta_forward expects one 3-channel panorama [1, 3, 512, 1024] (inference.py:196-200)')
    if x.shape[2] != PANO_H or x.shape[3] != PANO_W:
        raise NotImplementedError()
    dev = next(net.parameters()).device
    if dev.type != 'cuda':
        raise RuntimeError('horizonnet_b200 has no CPU path: move the model to a cuda device')
    shifts = [int(round(p * PANO_W)) for p in rotate]              # inference.py:40
    views = 1 + (1 if flip else 0) + len(shifts)
    h = net._handle(dev, max(views, 1))
    xd = x.detach().to(device=dev, dtype=torch.float32).contiguous()
    y_bon = torch.empty(2, PANO_W, device=dev, dtype=torch.float32)
    y_cor = torch.empty(PANO_W, device=dev, dtype=torch.float32)
    arr = (ctypes.c_int * max(len(shifts), 1))(*shifts) if shifts else None
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().hn_model_infer_tta(h['ptr'], xd.data_ptr(), 3, 1 if flip else 0, arr, len(shifts),
                                                 y_bon.data_ptr(), y_cor.data_ptr(), stream), 'hn_model_infer_tta')
    return y_bon.cpu().numpy(), y_cor.cpu().numpy()
