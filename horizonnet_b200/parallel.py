"""Data-parallel inference plumbing: panoramas are independent units, so a batch is cut
into contiguous shards, one process / GPU / library handle per shard, weights replicated, and the only
collective is one all-gather of the fused [B_local, 3, 1024] output (393 KB per rank at B_local = 32)
over NCCL (gloo in the CPU tests).  The reference has no multi-GPU inference (inference.py:184-188 is
single-device); training uses nn.DataParallel (train.py:190-192)."""
import torch
import torch.distributed as dist


def shard_bounds(total, rank, world):
    """[lo, hi) of rank's contiguous shard; the first `total % world` ranks get one extra unit."""
    if not (0 <= rank < world):
        raise ValueError('rank out of range')
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_outputs(bon, cor, group=None):
    """All-gathers (bon [b,2,1024], cor [b,1,1024]) from every rank; equal shard sizes required.
    Returns (bon_all [world*b,2,1024], cor_all [world*b,1,1024]) in rank order on every rank."""
    world = dist.get_world_size(group)
    fused = torch.cat([cor, bon], dim=1).contiguous()          # channel 0 = cor, 1..2 = bon (model.py:278-279)
    parts = [torch.empty_like(fused) for _ in range(world)]
    dist.all_gather(parts, fused, group=group)
    full = torch.cat(parts, dim=0)
    return full[:, 1:], full[:, :1]
