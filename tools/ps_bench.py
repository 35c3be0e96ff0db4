import sys, torch, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as e; e.build()
from horizonnet_b200.misc.panostretch import pano_stretch_batch
n=64
imgs=torch.rand(n,512,1024,3,device='cuda'); out=torch.empty_like(imgs)
kx=[0.5+1.5*i/n for i in range(n)]; ky=[2.0-1.5*i/n for i in range(n)]
for _ in range(3): pano_stretch_batch(imgs,kx,ky,out=out)
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): pano_stretch_batch(imgs,kx,ky,out=out)
b.record(); torch.cuda.synchronize()
ms=a.elapsed_time(b)/10
print(os.environ.get('HN_PS_F32'), 'ms', ms, 'GB/s', n*12582912/ms/1e6)
