"""Helpers for the -m gpu tests: torch is used only to build inputs in the library's halo-NHWC
layout and to hold device memory; every computation under test goes through the C ABI."""
import ctypes
import torch
import torch.nn.functional as F

from horizonnet_b200 import _lib


def to_halo_nhwc(x_nchw, halo):
    """NCHW -> [B, H, W+2*halo, C] with circular wrap columns."""
    x = x_nchw.permute(0, 2, 3, 1)
    if halo:
        x = torch.cat([x[:, :, -halo:], x, x[:, :, :halo]], dim=2)
    return x.contiguous()


def from_halo_nhwc(y, halo):
    if halo:
        y = y[:, :, halo:-halo]
    return y.permute(0, 3, 1, 2).contiguous()


def pack_weight(w_oihw):
    co, ci, kh, kw = w_oihw.shape
    return w_oihw.permute(2, 3, 1, 0).reshape(kh * kw * ci, co).contiguous()


def conv2d(x_nchw, w_oihw, scale, shift, stride, ph, pw, relu, residual_nchw=None, impl=0, in_halo=1, out_halo=1):
    """Runs hn_conv2d on cuda tensors, returns (NCHW interior, raw halo-NHWC output)."""
    lib = _lib.lib()
    dev = x_nchw.device
    B, Ci, H, W = x_nchw.shape
    Co, _, kh, kw = w_oihw.shape
    sh, sw = stride
    Ho = (H + 2 * ph - kh) // sh + 1
    Wo = (W + 2 * pw - kw) // sw + 1
    xin = to_halo_nhwc(x_nchw.float(), in_halo)
    wp = pack_weight(w_oihw.float())
    out = torch.full((B, Ho, Wo + 2 * out_halo, Co), float('nan'), device=dev)
    res = to_halo_nhwc(residual_nchw.float(), out_halo) if residual_nchw is not None else None
    stream = torch.cuda.current_stream(dev).cuda_stream
    rc = lib.hn_conv2d(xin.data_ptr(), B, H, W, Ci, in_halo, wp.data_ptr(), scale.contiguous().data_ptr(),
                       shift.contiguous().data_ptr(), res.data_ptr() if res is not None else None, Co, kh, kw, sh, sw,
                       ph, pw, int(relu), out.data_ptr(), out_halo, impl, stream)
    _lib.check(rc, 'hn_conv2d')
    torch.cuda.synchronize()
    return from_halo_nhwc(out, out_halo), out


def conv2d_reference(x, w, scale, shift, stride, ph, pw, relu, residual=None):
    """Plain PyTorch fp64 reference of the same op (circular W pad, zero H pad)."""
    x = x.double()
    if pw:
        x = torch.cat([x[..., -pw:], x, x[..., :pw]], dim=3)
    y = F.conv2d(x, w.double(), None, stride=stride, padding=(ph, 0))
    y = y * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)
    if residual is not None:
        y = y + residual.double()
    if relu:
        y = F.relu(y)
    return y
