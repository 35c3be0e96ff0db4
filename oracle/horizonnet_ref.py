"""TEST INFRASTRUCTURE ONLY -- CPU oracle for ``HorizonNet('resnet50', use_rnn=True).forward``.

A functional restatement, in plain CPU PyTorch ops, of the algorithm the reference runs
(reference = sunset1995/HorizonNet @ c9a7df9; every step cites the reference file:line it follows).
The arithmetic itself lives in third-party code that is not under /root/reference:
PyTorch (conv2d / batch_norm / interpolate; reference call sites model.py:129-131,154,222-230) and
torchvision's ``resnet50`` topology (model.py:66-68; torchvision/models/resnet.py Bottleneck v1.5).
The reference pins no torch version (environment.yml lists none; README.md:28 "pytorch 1.8.1");
the oracle is pinned to torch 2.11.0 CPU fp32 -- the only version available on any box.

Pinning: the reference ships NO tests or golden vectors for this path, so the oracle is pinned
against the reference itself, imported read-only in the build container:
tests/golden/make_golden.py loads the same synthetic checkpoint into the real ``model.HorizonNet``
and stores its outputs + per-stage samples in tests/golden/forward_*.npz;
tests/test_oracle.py asserts this restatement reproduces them.

It takes the reference's 448-key ``state_dict`` directly, so it also documents the checkpoint
layout (reference misc/utils.py:49-65).
"""
import torch
import torch.nn.functional as F

X_MEAN = (0.485, 0.456, 0.406)      # model.py:186
X_STD = (0.229, 0.224, 0.225)       # model.py:187
BN_EPS = 1e-5                       # torch.nn.BatchNorm2d default used by torchvision + model.py:130


def _circ_conv(x, w, b, stride, pad_h, pad_w):
    # model.py:27-29 (lr_pad) + model.py:42-55 (wrap_lr_pad): horizontal zero padding of every
    # conv with padding[1] != 0 is replaced by a left/right wrap; vertical zero padding is kept.
    if pad_w:
        x = torch.cat([x[..., -pad_w:], x, x[..., :pad_w]], dim=3)
    return F.conv2d(x, w, b, stride=stride, padding=(pad_h, 0))


class TrainMode:
    """What ``net.train()`` changes in the forward (train.py:52 runs it under net.train()):
      * every nn.BatchNorm2d not listed in ``frozen`` (state_dict prefixes; train.py:251-256 keeps the frozen blocks in
        eval mode) normalises with batch statistics and moves its running statistics by ``momentum`` (torch default 0.1;
        train.py:210-213 --bn_momentum) -- the moved values are collected in ``running``;
      * nn.LSTM(dropout=p) drops layer-1 outputs on their way into layer 2 (model.py:226) and self.drop_out drops the
        LSTM output before the linear head (model.py:228, :265).  ``masks`` = (inter-layer, head) multiplicative masks
        [256, B, 1024] (0 or 1/(1-p)); None draws them with F.dropout from torch's generator in that same order, which is
        the order the reference consumes it in, so torch.manual_seed reproduces the reference's own masks on CPU.
    """

    def __init__(self, masks=None, momentum=0.1, frozen=(), p=0.5, relu_masks=None):
        self.masks, self.momentum, self.frozen, self.p = masks, momentum, set(frozen), p
        self.running = {}
        # {BatchNorm2d prefix: bool tensor}: take the ReLU decisions behind that BatchNorm from the caller instead of
        # from the sign of this run's own pre-activations.  Two fp32 implementations of this forward differ by ~1e-4
        # deep in the net, so ~1 % of the near-zero pre-activations flip sign between them; a gradient comparison has
        # to factor those flips out (tests/test_gpu_parity.py) -- they are not errors of either side.
        self.relu_masks = relu_masks

    def dropout(self, x, which):
        if self.p <= 0:
            return x
        return x * self.masks[which].to(x.dtype) if self.masks is not None else F.dropout(x, self.p, True)


def _relu(x, p, tm=None):
    m = tm.relu_masks.get(p) if (tm is not None and tm.relu_masks) else None
    return F.relu(x) if m is None else x * m.to(x.dtype)


def _bn(x, sd, p, tm=None):
    if tm is not None and p not in tm.frozen:
        rm, rv = sd[p + '.running_mean'].clone(), sd[p + '.running_var'].clone()
        y = F.batch_norm(x, rm, rv, sd[p + '.weight'], sd[p + '.bias'], True, tm.momentum, BN_EPS)
        tm.running[p + '.running_mean'], tm.running[p + '.running_var'] = rm, rv
        return y
    return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'],
                        sd[p + '.weight'], sd[p + '.bias'], False, 0.0, BN_EPS)


def _bottleneck(x, sd, p, stride, has_ds, tm=None):
    # torchvision resnet.py Bottleneck.forward (v1.5: the stride sits on the 3x3 conv2)
    out = _relu(_bn(F.conv2d(x, sd[p + 'conv1.weight']), sd, p + 'bn1', tm), p + 'bn1', tm)
    out = _relu(_bn(_circ_conv(out, sd[p + 'conv2.1.weight'], None, stride, 1, 1), sd, p + 'bn2', tm), p + 'bn2', tm)
    out = _bn(F.conv2d(out, sd[p + 'conv3.weight']), sd, p + 'bn3', tm)
    if has_ds:
        x = _bn(F.conv2d(x, sd[p + 'downsample.0.weight'], stride=stride), sd, p + 'downsample.1', tm)
    return _relu(out + x, p + 'bn3', tm)


def encoder(x, sd, tm=None):
    """model.py:71-82 (Resnet.forward): stem + layer1..4, returns the 4 feature maps."""
    e = 'feature_extractor.encoder.'
    x = _circ_conv(x, sd[e + 'conv1.1.weight'], None, 2, 3, 3)        # model.py:73 (7x7 s2, wrapped)
    x = _relu(_bn(x, sd, e + 'bn1', tm), e + 'bn1', tm)                # model.py:74-75
    x = F.max_pool2d(x, 3, 2, 1)                                       # model.py:76 (NOT wrapped)
    feats = []
    for li, nblk in zip((1, 2, 3, 4), (3, 4, 6, 3)):
        for b in range(nblk):
            stride = 2 if (b == 0 and li > 1) else 1
            x = _bottleneck(x, sd, f'{e}layer{li}.{b}.', stride, b == 0, tm)
        feats.append(x)                                                # model.py:78-81
    return feats


def global_height_conv(x, sd, s, out_w, tm=None):
    """model.py:148-156 (GlobalHeightConv.forward) for scale index s."""
    for j in range(4):
        p = f'reduce_height_module.ghc_lst.{s}.layer.{j}.layers.'
        x = _circ_conv(x, sd[p + '0.1.weight'], sd[p + '0.1.bias'], (2, 1), 1, 1)   # model.py:129
        x = _relu(_bn(x, sd, p + '1', tm), p + '1', tm)                              # model.py:130-131
    factor = out_w // x.shape[3]                                                     # model.py:152
    x = torch.cat([x[..., -1:], x, x[..., :1]], 3)                                   # model.py:153
    x = F.interpolate(x, size=(x.shape[2], out_w + 2 * factor), mode='bilinear',
                      align_corners=False)                                           # model.py:154
    return x[..., factor:-factor]                                                    # model.py:155


def lstm_layer_dir(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of one nn.LSTM layer (model.py:222-227), h0 = c0 = 0, gate order i,f,g,o."""
    T, B, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    xp = x @ w_ih.t() + (b_ih + b_hh)
    out = x.new_empty(T, B, H)
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        g = xp[t] + h @ w_hh.t()
        i, f, gg, o = g.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[t] = h
    return out


def bi_lstm(x, sd, tm=None):
    for layer in range(2):
        if layer == 1 and tm is not None:
            x = tm.dropout(x, 0)          # nn.LSTM(dropout=0.5), train mode only (model.py:226)
        outs = []
        for suffix, rev in (('', False), ('_reverse', True)):
            outs.append(lstm_layer_dir(
                x, sd[f'bi_rnn.weight_ih_l{layer}{suffix}'], sd[f'bi_rnn.weight_hh_l{layer}{suffix}'],
                sd[f'bi_rnn.bias_ih_l{layer}{suffix}'], sd[f'bi_rnn.bias_hh_l{layer}{suffix}'], rev))
        x = torch.cat(outs, dim=2)
    return x


def forward(sd, x, dtype=torch.float32, return_stages=False, train=None):
    """model.py:254-281.  sd: the 448-key state_dict; x: [B, C>=3, 512, 1024] in [0,1].

    Returns (bon [B,2,1024], cor [B,1,1024]) like the reference; with return_stages also a dict of
    intermediate tensors (NCHW) used by the per-stage parity tests.  train: a TrainMode (train-mode forward; the
    moved running statistics end up in train.running), None = eval.
    """
    tm = train
    if x.shape[2] != 512 or x.shape[3] != 1024:
        raise NotImplementedError()                                    # model.py:255-256
    sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    x = x.to(dtype)
    mean = x.new_tensor(X_MEAN).view(1, 3, 1, 1)
    std = x.new_tensor(X_STD).view(1, 3, 1, 1)
    x = (x[:, :3] - mean) / std                                        # model.py:248-252
    feats = encoder(x, sd, tm)
    bs = x.shape[0]
    out_w = 1024 // 4                                                  # model.py:260
    red = [global_height_conv(f, sd, s, out_w, tm).reshape(bs, -1, out_w) for s, f in enumerate(feats)]
    feature = torch.cat(red, dim=1)                                    # model.py:175-178 -> [B,1024,256]
    seq = feature.permute(2, 0, 1)                                     # model.py:263
    rnn_out = bi_lstm(seq, sd, tm)                                     # model.py:264 (dropout = id in eval)
    head_in = rnn_out if tm is None else tm.dropout(rnn_out, 1)        # model.py:265 self.drop_out
    out = head_in @ sd['linear.weight'].t() + sd['linear.bias']        # model.py:266
    out = out.view(out.shape[0], out.shape[1], 3, 4).permute(1, 2, 0, 3)
    out = out.contiguous().view(out.shape[0], 3, -1)                   # model.py:267-269
    cor, bon = out[:, :1], out[:, 1:]                                  # model.py:278-279
    if return_stages:
        stages = {f'layer{i + 1}': f for i, f in enumerate(feats)}
        stages['feature'] = feature
        stages['rnn_out'] = rnn_out
        return bon, cor, stages
    return bon, cor
