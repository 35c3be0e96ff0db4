"""Loader of tests/golden/train_*.npz (minted by make_golden.py golden_train from the REAL reference under net.train())."""
import os
import numpy as np
import torch

from horizonnet_b200.weights import synthetic_state_dict, synthetic_panoramas


def train_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f'train_{name}.npz'))
    B = int(g['batch'])
    bits = np.unpackbits(g['masks_bits'])[:2 * 256 * B * 1024].reshape(2, 256, B, 1024)
    masks = [torch.from_numpy(bits[i].astype(np.float32) * 2.0) for i in range(2)]
    sd = synthetic_state_dict(int(g['wseed']), 'random')
    x = synthetic_panoramas(B, seed=int(g['x_seed']))
    bn_keys = [k for k in sd if k.endswith('running_mean') or k.endswith('running_var')]
    running, off = {}, 0
    for k in bn_keys:
        n = sd[k].numel()
        running[k] = torch.from_numpy(g['running'][off:off + n].copy())
        off += n
    return g, sd, x, masks, running, [str(f) for f in g['frozen']]
