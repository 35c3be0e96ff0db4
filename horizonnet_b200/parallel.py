"""Data-parallel inference plumbing: panoramas are independent units, so a batch is cut
into contiguous shards, one process / GPU / library handle per shard, weights replicated, and the only
collective is one all-gather of the outputs (393 KB per rank at B_local = 32) over NCCL / NVLink (gloo in the
CPU tests).  The reference has no multi-GPU inference (inference.py:184-188 is single-device); training uses
nn.DataParallel (train.py:190-192), i.e. per-GPU BatchNorm statistics and a gradient sum on GPU 0 -- here: one
process per GPU, per-GPU BatchNorm statistics, and ``average_gradients`` (bucketed all-reduce) between
``loss.backward()`` and ``optimizer.step()``."""
import torch
import torch.distributed as dist


def shard_bounds(total, rank, world):
    """[lo, hi) of rank's contiguous shard; the first `total % world` ranks get one extra unit."""
    if not (0 <= rank < world):
        raise ValueError('rank out of range')
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class OutputGatherer:
    """All-gather of (bon [b,2,1024], cor [b,1,1024]) into preallocated [world*b, ...] buffers with
    ``all_gather_into_tensor`` (no list API, no concatenation kernels): NCCL writes every rank's shard straight
    into its slot of the result, in rank order.  Two rotating result buffers, so that the gather of step i can
    still be in flight while step i+1 is being enqueued.  With ``total`` given, unequal shards
    (``shard_bounds(total, rank, world)``) are padded to the largest shard for the collective and trimmed after."""

    def __init__(self, group=None, total=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.total = total
        self._bufs = {}
        self._turn = 0

    def _buffers(self, b, device):
        key = (b, str(device))
        if key not in self._bufs:
            self._bufs[key] = [(torch.empty(self.world * b, 2, 1024, device=device, dtype=torch.float32),
                                torch.empty(self.world * b, 1, 1024, device=device, dtype=torch.float32))
                               for _ in range(2)]
        self._turn ^= 1
        return self._bufs[key][self._turn]

    def __call__(self, bon, cor):
        b = bon.shape[0]
        sizes = None
        if self.total is not None:
            sizes = [hi - lo for lo, hi in (shard_bounds(self.total, r, self.world) for r in range(self.world))]
            if sizes[self.rank] != b:
                raise ValueError(f'rank {self.rank} holds {b} panoramas, shard_bounds says {sizes[self.rank]}')
            bmax = max(sizes)
            if b < bmax:        # pad to the largest shard: the collective needs equal shapes on every rank
                bon = torch.cat([bon, bon.new_zeros(bmax - b, 2, 1024)], dim=0)
                cor = torch.cat([cor, cor.new_zeros(bmax - b, 1, 1024)], dim=0)
            b = bmax
        bon_all, cor_all = self._buffers(b, bon.device)
        dist.all_gather_into_tensor(bon_all, bon.contiguous(), group=self.group)
        dist.all_gather_into_tensor(cor_all, cor.contiguous(), group=self.group)
        if sizes is not None and any(s != b for s in sizes):
            keep = torch.cat([torch.arange(r * b, r * b + s) for r, s in enumerate(sizes)]).to(bon_all.device)
            return bon_all.index_select(0, keep), cor_all.index_select(0, keep)
        return bon_all, cor_all


_default = {}


def gather_outputs(bon, cor, group=None, total=None):
    """All-gathers (bon [b,2,1024], cor [b,1,1024]) from every rank.  Returns (bon_all, cor_all) in rank order on
    every rank.  Equal shard sizes unless ``total`` is given (then shards follow ``shard_bounds``)."""
    key = (id(group), total)
    if key not in _default:
        _default[key] = OutputGatherer(group, total)
    return _default[key](bon, cor)


def average_gradients(params, group=None, bucket_bytes=64 << 20):
    """Data-parallel training (replaces nn.DataParallel's gradient reduction, train.py:190-192): averages ``.grad`` of
    the given parameters over the ranks with all-reduces of flat buckets of about ``bucket_bytes`` (326 MB of fp32
    gradients for HorizonNet resnet50+rnn => 6 collectives; NCCL rings / NVLS are bandwidth-bound at that size).
    Every rank must pass the parameters in the same order.  Parameters without a gradient are skipped on the
    condition that they have none on every rank (frozen blocks, train.py:200-208).  Returns the number of
    collectives issued."""
    world = dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    if world == 1 or not grads:
        return 0
    calls = 0
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size, calls
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        off = 0
        for g in bucket:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
        calls += 1
        bucket, size = [], 0

    for g in grads:
        if bucket and (g.dtype != bucket[0].dtype or g.device != bucket[0].device):
            flush()
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
    return calls
