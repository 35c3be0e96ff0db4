"""One pass over every entry point of the library at small sizes, for compute-sanitizer
(memcheck / racecheck / synccheck / initcheck).  Run: tools/sanitize.sh"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as e; e.build()
from horizonnet_b200.model import HorizonNet
from horizonnet_b200.weights import synthetic_state_dict, synthetic_panoramas
from horizonnet_b200.misc.panostretch import pano_stretch, pano_stretch_batch
from horizonnet_b200.misc.pano_lsd_align import rotatePanorama
from horizonnet_b200 import augment as aug

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
what = sys.argv[2] if len(sys.argv) > 2 else 'all'
dev = 'cuda:0'
if what in ('all', 'model'):
    net = HorizonNet('resnet50', True).eval(); net.load_state_dict(synthetic_state_dict(0, 'random')); net = net.to(dev)
    x = synthetic_panoramas(bs, seed=1000).to(dev)
    with torch.no_grad():
        b0, c0 = net(x)                               # plain forward (fused layer1 + tensor-core convs + cluster LSTM)
        net.forward_pipelined(x); net.forward_pipelined(x); outs = net.flush()   # two-stream schedule
    torch.cuda.synchronize()
    assert torch.equal(outs[-1][0], b0) and torch.equal(outs[-1][1], c0)
    print('model ok', float(b0.abs().max()))
if what in ('all', 'aux'):
    rng = np.random.RandomState(0)
    img = rng.rand(64, 128, 3).astype(np.float32)
    o, _ = pano_stretch(img, np.array([[10., 20.]]), 1.3, 0.8)
    o64, _ = pano_stretch(img.astype(np.float64), np.array([[10., 20.]]), 0.7, 1.6, order=0)
    t = torch.from_numpy(rng.rand(3, 64, 128, 3).astype(np.float32)).to(dev)
    pano_stretch_batch(t, [1.1, 0.9, 1.5], [0.6, 1.9, 1.0])
    u8 = torch.from_numpy(rng.randint(0, 256, (3, 64, 128, 3), dtype=np.uint8)).to(dev)
    aug.augment_batch(u8, kx=[1.2, 1.0, 0.7], ky=[0.8, 1.0, 1.4], flip=[1, 0, 1], dx=[5, 0, 127], gamma=[1.5, 1.0, 0.6])
    a = 0.3; R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    rotatePanorama(rng.rand(64, 128, 3), R=R)
    torch.cuda.synchronize()
    print('aux ok')
