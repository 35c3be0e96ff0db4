"""Data-parallel training step under torchrun (one process per GPU, NCCL): train.py's nn.DataParallel step (train.py:190-192,
272-281) as forward + backward through the library, parallel.average_gradients (bucketed all-reduce), Adam.
Checks that every rank ends up with the mean of the per-rank gradients; prints one JSON line on rank 0.
Usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/train_ddp_bench.py [batch] [steps]"""
import json, os, sys
import torch, torch.distributed as dist, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as e
from horizonnet_b200.model import HorizonNet
from horizonnet_b200.parallel import average_gradients
from horizonnet_b200.weights import synthetic_state_dict, synthetic_panoramas

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
if rank == 0:
    e.build()
dist.barrier()
net = HorizonNet('resnet50', True)
net.load_state_dict(synthetic_state_dict(0, 'random'))
net = net.to(dev).train()
params = list(net.parameters())
opt = torch.optim.Adam(params, lr=1e-4, betas=(0.9, 0.999))
x = synthetic_panoramas(B, seed=100 + rank).to(dev)                  # every rank its own shard of the global batch
g = torch.Generator().manual_seed(200 + rank)
y_bon, y_cor = (torch.rand(B, 2, 1024, generator=g) - 0.5).to(dev), torch.rand(B, 1, 1024, generator=g).to(dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
acc = [0.0] * 4
check = None
for it in range(steps + 2):
    dist.barrier(); torch.cuda.synchronize(dev)
    ev[0].record()
    opt.zero_grad()
    bon, cor = net(x)
    loss = F.l1_loss(bon, y_bon) + F.binary_cross_entropy_with_logits(cor, y_cor)
    ev[1].record()
    loss.backward()
    ev[2].record()
    if it == 0:                     # the mean of the per-rank gradients of one tensor, the slow way, as the checker
        local_g = net.linear.weight.grad.clone()
        allg = [torch.empty_like(local_g) for _ in range(world)]
        dist.all_gather(allg, local_g)
        want = sum(allg) / world
    calls = average_gradients(params)
    ev[3].record()
    if it == 0:
        check = bool(torch.allclose(net.linear.weight.grad, want, rtol=1e-6, atol=1e-9))
    opt.step()
    ev[4].record()
    torch.cuda.synchronize(dev)
    if it >= 2:
        for i in range(4):
            acc[i] += ev[i].elapsed_time(ev[i + 1])
net.check()
t = torch.tensor([a / steps for a in acc], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)                            # max over ranks, per phase
fw, bw, ar, op = t.tolist()
ok = torch.tensor([1 if check else 0], device=dev)
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    nbytes = sum(p.numel() * 4 for p in params)
    print(json.dumps({'n_gpus': world, 'batch_per_gpu': B, 'forward_ms': round(fw, 2), 'backward_ms': round(bw, 2),
                      'grad_allreduce_ms': round(ar, 2), 'optimizer_ms': round(op, 2), 'step_ms': round(fw + bw + ar + op, 2),
                      'panoramas_per_s': round(world * B / (fw + bw + ar + op) * 1e3, 2), 'allreduce_collectives': calls,
                      'gradient_bytes': nbytes, 'allreduce_busbw_gbs': round(2 * (world - 1) / world * nbytes / (ar * 1e-3) / 1e9, 1),
                      'averaged_gradient_equals_mean_of_ranks': bool(ok.item()), 'timing': 'CUDA events, max over ranks per phase'}))
dist.destroy_process_group()
