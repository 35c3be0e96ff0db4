"""Architecture spec of the one model family on the hot path: ``resnet50`` + bi-LSTM head.

Everything here is derived from the *checkpoint layout* the reference writes
(reference misc/utils.py:49-58, model.py:189-246) -- 448 keys, probed from the reference and
pinned by tests/golden/state_dict_keys.json.  The C library builds the same graph from the same
key names (csrc/model.cu), so this file is the single Python-side description of names + shapes.
"""
from collections import OrderedDict

RESNET50_BLOCKS = (3, 4, 6, 3)
RESNET50_PLANES = (64, 128, 256, 512)
EXPANSION = 4
PANO_H, PANO_W = 512, 1024          # reference model.py:255 hard-wires the input size
STEP_COLS = 4                        # model.py:194
SEQ_LEN = PANO_W // STEP_COLS        # 256 LSTM steps
RNN_HIDDEN = 512                     # model.py:195
OUT_SCALE = 8                        # model.py:193
HEAD_BIAS = [-1.0] * 4 + [-0.478] * 4 + [0.425] * 4   # model.py:231-233


def _bn(prefix, c, out):
    out[prefix + '.weight'] = ((c,), 'bn_weight')
    out[prefix + '.bias'] = ((c,), 'bn_bias')
    out[prefix + '.running_mean'] = ((c,), 'bn_mean')
    out[prefix + '.running_var'] = ((c,), 'bn_var')
    out[prefix + '.num_batches_tracked'] = ((), 'bn_count')


def state_dict_spec():
    """OrderedDict key -> (shape, kind) in the reference's own key order."""
    out = OrderedDict()
    enc = 'feature_extractor.encoder.'
    out[enc + 'conv1.1.weight'] = ((64, 3, 7, 7), 'enc_conv')       # wrapped by LR_PAD -> '.1'
    _bn(enc + 'bn1', 64, out)
    inplanes = 64
    for li, (nblk, planes) in enumerate(zip(RESNET50_BLOCKS, RESNET50_PLANES), start=1):
        for b in range(nblk):
            p = f'{enc}layer{li}.{b}.'
            out[p + 'conv1.weight'] = ((planes, inplanes, 1, 1), 'enc_conv')
            _bn(p + 'bn1', planes, out)
            out[p + 'conv2.1.weight'] = ((planes, planes, 3, 3), 'enc_conv')   # wrapped 3x3
            _bn(p + 'bn2', planes, out)
            out[p + 'conv3.weight'] = ((planes * EXPANSION, planes, 1, 1), 'enc_conv')
            _bn(p + 'bn3', planes * EXPANSION, out)
            if b == 0:
                out[p + 'downsample.0.weight'] = ((planes * EXPANSION, inplanes, 1, 1), 'enc_conv')
                _bn(p + 'downsample.1', planes * EXPANSION, out)
            inplanes = planes * EXPANSION
    for s, c in enumerate(encoder_channels()):
        chans = [c, c // 2, c // 2, c // 4, c // OUT_SCALE]
        for j in range(4):
            p = f'reduce_height_module.ghc_lst.{s}.layer.{j}.layers.'
            out[p + '0.1.weight'] = ((chans[j + 1], chans[j], 3, 3), 'ghc_conv')
            out[p + '0.1.bias'] = ((chans[j + 1],), 'ghc_bias')
            _bn(p + '1', chans[j + 1], out)
    c_last = sum(c * h for c, h in zip(encoder_channels(), (8, 4, 2, 1))) // OUT_SCALE   # 1024
    for layer in range(2):
        for suffix in ('', '_reverse'):
            in_sz = c_last if layer == 0 else 2 * RNN_HIDDEN
            out[f'bi_rnn.weight_ih_l{layer}{suffix}'] = ((4 * RNN_HIDDEN, in_sz), 'rnn')
            out[f'bi_rnn.weight_hh_l{layer}{suffix}'] = ((4 * RNN_HIDDEN, RNN_HIDDEN), 'rnn')
            out[f'bi_rnn.bias_ih_l{layer}{suffix}'] = ((4 * RNN_HIDDEN,), 'rnn')
            out[f'bi_rnn.bias_hh_l{layer}{suffix}'] = ((4 * RNN_HIDDEN,), 'rnn')
    out['linear.weight'] = ((3 * STEP_COLS, 2 * RNN_HIDDEN), 'head_weight')
    out['linear.bias'] = ((3 * STEP_COLS,), 'head_bias')
    return out


def encoder_channels():
    return tuple(p * EXPANSION for p in RESNET50_PLANES)     # 256, 512, 1024, 2048
