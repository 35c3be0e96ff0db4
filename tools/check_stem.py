"""Debug aid: tcgen05 stem (stem_tc_kernel) vs the oracle stem in fp64, with an error breakdown.
usage (GPU box): python tools/check_stem.py [batch]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import __graft_entry__ as entry
from horizonnet_b200.model import HorizonNet
from horizonnet_b200.weights import synthetic_state_dict, synthetic_panoramas
from oracle import horizonnet_ref as R

entry.build()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sd = synthetic_state_dict(5, 'random')
net = HorizonNet('resnet50', True).eval(); net.load_state_dict(sd); net.use_tensor_cores(True); net = net.to('cuda:0')
x = synthetic_panoramas(B, seed=33)
e = 'feature_extractor.encoder.'
sd64 = {k: v.double() for k, v in sd.items() if k.startswith(e + 'conv1') or k.startswith(e + 'bn1.')}
xn = (x[:, :3].double() - torch.tensor(R.X_MEAN).double().view(1, 3, 1, 1)) / torch.tensor(R.X_STD).double().view(1, 3, 1, 1)
pre = R._bn(R._circ_conv(xn, sd64[e + 'conv1.1.weight'], None, 2, 3, 3), sd64, e + 'bn1')
ref = F.relu(pre)
print('ref maxabs', ref.abs().max().item(), 'mean', ref.mean().item())
for mode in (1, 0):
    with torch.no_grad():
        net(x.to('cuda:0'))
        net.set_option('stem_tc', mode)
        bon, cor = net(x.to('cuda:0'))
    try:
        net.check()
    except Exception as ex:
        print('check failed', ex)
    st = net.debug_stage('stem').cpu().double()
    d = (st - ref).abs()
    print(f'mode {mode}: max err {d.max().item():.3e}  mean err {d.mean().item():.3e}  got maxabs {st.abs().max().item():.4f}')
    if d.max().item() > 1e-4:
        print('  err by channel (first 8):', d.amax(dim=(0, 2, 3))[:8].tolist())
        print('  err by x mod 8:', [d[..., i::8].max().item() for i in range(8)])
        print('  err by x segment:', [d[..., i * 128:(i + 1) * 128].max().item() for i in range(4)])
        print('  err by row (first 6):', d.amax(dim=(0, 1, 3))[:6].tolist(), 'last 3', d.amax(dim=(0, 1, 3))[-3:].tolist())
        print('  err by image:', d.amax(dim=(1, 2, 3)).tolist())
        print('  sample got', st[0, :4, 10, 10].tolist(), 'ref', ref[0, :4, 10, 10].tolist())
        ratio = (st[0, :, 8:200, 8:200] / ref[0, :, 8:200, 8:200].clamp_min(1e-3))
        print('  median ratio got/ref where ref>1e-3:', ratio[ref[0, :, 8:200, 8:200] > 1e-3].median().item())
    print(f'  bon maxabs {bon.abs().max().item():.4f}')
