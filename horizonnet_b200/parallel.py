"""Data-parallel inference plumbing: panoramas are independent units, so a batch is cut
into contiguous shards, one process / GPU / library handle per shard, weights replicated, and the only
collective is one all-gather of the outputs (393 KB per rank at B_local = 32) over NCCL / NVLink (gloo in the
CPU tests).  The reference has no multi-GPU inference (inference.py:184-188 is single-device); training uses
nn.DataParallel (train.py:190-192)."""
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
