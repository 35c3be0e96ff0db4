"""Device time of one training step (BASELINE config 5: batch 8 per GPU, L1 + BCE-with-logits, Adam) through the library:
forward with tape, backward, gradient read-back, optimizer step, weight re-upload.  Usage: tools/train_bench.py [batch] [steps]"""
import os, sys, time, json, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as e; e.build()
from horizonnet_b200.model import HorizonNet
from horizonnet_b200.weights import synthetic_state_dict, synthetic_panoramas
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = 'cuda:0'
net = HorizonNet('resnet50', True); net.load_state_dict(synthetic_state_dict(0, 'random')); net = net.to(dev).train()
opt = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.9, 0.999))
x = synthetic_panoramas(B, seed=5).to(dev)
y_bon = (torch.rand(B, 2, 1024) - 0.5).to(dev); y_cor = torch.rand(B, 1, 1024).to(dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
res = []
for it in range(steps + 2):
    torch.cuda.synchronize(); ev[0].record()
    opt.zero_grad()
    bon, cor = net(x)
    loss = F.l1_loss(bon, y_bon) + F.binary_cross_entropy_with_logits(cor, y_cor)
    ev[1].record()
    loss.backward()
    ev[2].record()
    opt.step()
    ev[3].record()
    torch.cuda.synchronize()
    if it >= 2:
        res.append((ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])))
fw, bw, op = (sum(r[i] for r in res) / len(res) for i in range(3))
print(json.dumps({'batch': B, 'forward_ms': round(fw, 2), 'backward_ms': round(bw, 2), 'adam_ms': round(op, 2),
                  'step_ms': round(fw + bw + op, 2), 'panoramas_per_s': round(B / (fw + bw + op) * 1e3, 2),
                  'loss': loss.item(), 'mem_gb': round(torch.cuda.max_memory_allocated() / 2**30, 2)}))
