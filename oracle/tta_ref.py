"""TEST INFRASTRUCTURE ONLY -- numpy/torch restatement of the TTA part of reference inference.py.

Follows inference.py:32-43 (augment), :46-62 (augment_undo) and :77-93 (forward on the views, sigmoid,
un-augment, mean, boundary -> pixel rows, clipping), with the oracle forward (oracle/horizonnet_ref.py) in
place of ``net``.  Pinned by tests/golden/tta_identity.npz (minted from the real reference's
``inference(..., force_raw=True)`` internals by tests/golden/make_golden.py)."""
import numpy as np
import torch

from . import horizonnet_ref


def augment(x_img, flip, rotate):
    x_img = x_img.numpy()
    aug_type = ['']
    out = [x_img]
    if flip:
        aug_type.append('flip')
        out.append(np.flip(x_img, axis=-1))                                  # inference.py:36-38
    for shift_p in rotate:
        shift = int(round(shift_p * x_img.shape[-1]))                        # :40
        aug_type.append('rotate %d' % shift)
        out.append(np.roll(x_img, shift, axis=-1))                           # :42
    return torch.FloatTensor(np.concatenate(out, 0)), aug_type


def augment_undo(x_aug, aug_type):
    x_aug = x_aug.cpu().numpy()
    sz = x_aug.shape[0] // len(aug_type)
    outs = []
    for i, aug in enumerate(aug_type):
        part = x_aug[i * sz:(i + 1) * sz]
        if aug == 'flip':
            outs.append(np.flip(part, axis=-1))                              # :53-54
        elif aug.startswith('rotate'):
            outs.append(np.roll(part, -int(aug.split()[-1]), axis=-1))       # :55-57
        else:
            outs.append(part)
    return np.array(outs)


def tta_forward(sd, x, flip=False, rotate=()):
    H = x.shape[2]
    xa, aug_type = augment(x, flip, list(rotate))
    with torch.no_grad():
        y_bon_, y_cor_ = horizonnet_ref.forward(sd, xa)                      # :78
    y_bon_ = augment_undo(y_bon_, aug_type).mean(0)                          # :79
    y_cor_ = augment_undo(torch.sigmoid(y_cor_), aug_type).mean(0)           # :80
    y_bon_ = (y_bon_[0] / np.pi + 0.5) * H - 0.5                             # :90
    y_bon_[0] = np.clip(y_bon_[0], 1, H / 2 - 1)                             # :91
    y_bon_[1] = np.clip(y_bon_[1], H / 2 + 1, H - 2)                         # :92
    return y_bon_, y_cor_[0, 0]                                              # :93
