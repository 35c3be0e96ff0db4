"""GPU probe: how does tcgen05 kind::f16 accumulate into fp32 TMEM?  (run under gpurun)
1x1 conv = GEMM with fp16-exact operands (lo planes are zero), compare with fp64."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
import __graft_entry__ as entry
entry.build()
import gpu_utils as gu

dev = 'cuda:0'
res = []
for sign in ('pos', 'mixed'):
    for K in (64, 256, 1024, 4096):
        g = torch.Generator().manual_seed(K)
        # fp16-exact values: integers/64 in [0, 4)
        x = torch.randint(0, 256, (1, K, 8, 32), generator=g).float() / 256
        w = torch.randint(1 if sign == 'pos' else -255, 256, (128, K, 1, 1), generator=g).float() / 64
        scale = torch.ones(128); shift = torch.zeros(128)
        ref = gu.conv2d_reference(x, w, scale, shift, (1, 1), 0, 0, False)
        for impl in (0, 1):
            y, _ = gu.conv2d(x.to(dev), w.to(dev), scale.to(dev), shift.to(dev), (1, 1), 0, 0, False, impl=impl)
            rel = ((y.cpu().double() - ref) / ref.abs().clamp_min(1e-30))
            m = ref.abs() > 0.1 * ref.abs().max()
            res.append(dict(sign=sign, K=K, impl=impl, mean_rel=float(rel[m].mean()), max_rel=float(rel[m].abs().max()),
                            mma_steps=K // 16))
            print(res[-1])
json.dump(res, open('gpurun_out/tc_accum_probe.json', 'w'), indent=1)
