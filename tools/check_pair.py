"""GPU: exercise the CTA-pair (cta_group::2) variant of conv_tc_kernel (HN_TC_PAIR=1) against torch fp64."""
import os, sys
os.environ['HN_TC_PAIR'] = '1'
os.environ['HN_TC_ROWBOX'] = '0'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import __graft_entry__ as entry
entry.build()
import gpu_utils as gu

CASES = [(2, 512, 8, 32, 256, 3, (1, 1), False, True), (1, 512, 9, 62, 256, 1, (1, 1), False, True),
         (3, 256, 16, 64, 128, 3, (2, 1), False, True), (1, 1024, 4, 32, 4096, 1, (1, 1), False, False)]
worst = 0.0
for case in CASES:
    B, Ci, H, W, Co, k, stride, residual, relu = case
    g = torch.Generator().manual_seed(Ci + Co)
    x = torch.randn(B, Ci, H, W, generator=g)
    w = torch.randn(Co, Ci, k, k, generator=g) / (Ci * k * k) ** 0.5
    scale = torch.rand(Co, generator=g) + 0.5
    shift = torch.randn(Co, generator=g) * 0.1
    ref = gu.conv2d_reference(x, w, scale, shift, stride, k // 2, k // 2, relu, None)
    y, raw = gu.conv2d(x.cuda(), w.cuda(), scale.cuda(), shift.cuda(), stride, k // 2, k // 2, relu, None, impl=1)
    err = (y.cpu().double() - ref).abs().max().item() / max(1.0, ref.abs().max().item())
    assert torch.equal(raw[:, :, 0], raw[:, :, -2]) and torch.equal(raw[:, :, -1], raw[:, :, 1])
    print(case, 'rel err %.3e' % err)
    worst = max(worst, err)
assert worst < 1e-4, worst
print('PAIR OK')
