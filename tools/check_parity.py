"""GPU: print end-to-end max-abs errors of the CUDA path against the CPU oracle (run under gpurun)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry
entry.build()
from horizonnet_b200.model import HorizonNet
from horizonnet_b200.weights import synthetic_state_dict, synthetic_panoramas
from oracle import horizonnet_ref

out = []
for seed, bn, xseed in ((7, 'random', 21), (1, 'random', 101), (0, 'identity', 100), (3, 'random', 5)):
    sd = synthetic_state_dict(seed, bn)
    x = synthetic_panoramas(1, seed=xseed)
    with torch.no_grad():
        rb, rc = horizonnet_ref.forward(sd, x)
        b64, c64 = horizonnet_ref.forward(sd, x, dtype=torch.float64)
    for tc in (False, True):
        net = HorizonNet('resnet50', True).eval()
        net.load_state_dict(sd)
        net.use_tensor_cores(tc)
        net = net.to('cuda:0')
        with torch.no_grad():
            b, c = net(x.cuda())
        b, c = b.cpu(), c.cpu()
        net.check()
        rec = dict(seed=seed, bn=bn, tensor_cores=tc,
                   bon_vs_fp32=float((b - rb).abs().max()), cor_vs_fp32=float((c - rc).abs().max()),
                   bon_vs_fp64=float((b.double() - b64).abs().max()), cor_vs_fp64=float((c.double() - c64).abs().max()),
                   oracle32_vs_fp64=float((rb.double() - b64).abs().max()))
        print(rec)
        out.append(rec)
        del net
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'parity_%s.json' % os.environ.get('HN_TC_SEG', 'default')), 'w'), indent=1)
