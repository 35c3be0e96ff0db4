"""Quick device-timed comparison of library options (one process): plain forward bs32, stage split."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as e; e.build()
from horizonnet_b200.model import HorizonNet
from horizonnet_b200.weights import synthetic_state_dict, synthetic_panoramas
sd = synthetic_state_dict(0, 'random')
net = HorizonNet('resnet50', True).eval(); net.load_state_dict(sd); net = net.to('cuda:0')
xs = [synthetic_panoramas(32, seed=1000 + 100 * i).to('cuda:0') for i in range(2)]
def run(label, steps=10, pipelined=False, **opts):
    with torch.no_grad():
        net(xs[0])
        for k, v in opts.items(): net.set_option(k, v)
        for i in range(3): net(xs[i & 1])
        net.set_profile(not pipelined); net.read_profile(reset=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            (net.forward_pipelined if pipelined else net)(xs[i & 1])
        if pipelined: net.flush()
        b.record(); torch.cuda.synchronize()
        prof = net.read_profile(reset=True) if not pipelined else {}
        net.set_profile(False)
    print(json.dumps({'label': label, 'ms_per_step': round(a.elapsed_time(b) / steps, 3),
                      'stages': {k: round(v[0] / steps, 3) for k, v in prof.items()}}))
run('fused', fuse_bottleneck=1)
run('unfused', fuse_bottleneck=0)
run('fused', fuse_bottleneck=1)
run('fused-pipelined', pipelined=True, fuse_bottleneck=1)
run('unfused-pipelined', pipelined=True, fuse_bottleneck=0)
