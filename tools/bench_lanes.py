"""Throughput with L independent lanes (library handles) fed round-robin: kernels of different lanes fill each other's tails."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as e; e.build()
from horizonnet_b200.model import HorizonNet
from horizonnet_b200.weights import synthetic_state_dict, synthetic_panoramas
sd = synthetic_state_dict(0, 'random')
def mk():
    n = HorizonNet('resnet50', True).eval(); n.load_state_dict(sd); return n.to('cuda:0')
nets = [mk() for _ in range(3)]
xs = [synthetic_panoramas(32, seed=1000 + 100 * i).to('cuda:0') for i in range(2)]
def run(lanes, steps=24):
    with torch.no_grad():
        for i in range(2 * lanes): nets[i % lanes].forward_pipelined(xs[i & 1])
        for n in nets[:lanes]: n.flush()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        outs = [nets[i % lanes].forward_pipelined(xs[i & 1]) for i in range(steps)]
        for n in nets[:lanes]: n.flush()
        b.record(); torch.cuda.synchronize()
        ref = nets[0](xs[(steps - 1) & 1])
        ok = torch.equal(outs[-1][0], ref[0]) and torch.equal(outs[-1][1], ref[1])
    print(json.dumps({'lanes': lanes, 'ms_per_step': round(a.elapsed_time(b) / steps, 3), 'bitwise_ok': ok}))
for lanes in (1, 2, 3, 1, 2):
    run(lanes)
