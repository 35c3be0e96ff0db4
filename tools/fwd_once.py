"""A few plain bs32 forwards (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as e; e.build()
from horizonnet_b200.model import HorizonNet
from horizonnet_b200.weights import synthetic_state_dict, synthetic_panoramas
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
net = HorizonNet('resnet50', True).eval(); net.load_state_dict(synthetic_state_dict(0, 'random')); net = net.to('cuda:0')
x = synthetic_panoramas(32, seed=1000).to('cuda:0')
with torch.no_grad():
    for _ in range(n):
        net(x)
torch.cuda.synchronize()
print('done')
