"""CPU, world_size 2, gloo: the N>1 host logic (sharding + the output all-gather used by bench.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from horizonnet_b200.parallel import average_gradients, gather_outputs, shard_bounds


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(123)
        bon_full = torch.randn(total, 2, 1024, generator=g)      # the "single-device" result
        cor_full = torch.randn(total, 1, 1024, generator=g)
        lo, hi = shard_bounds(total, rank, world)
        bon_all, cor_all = gather_outputs(bon_full[lo:hi].clone(), cor_full[lo:hi].clone())
        # gathered == single-device result, bit for bit
        assert torch.equal(bon_all, bon_full) and torch.equal(cor_all, cor_full)
    finally:
        dist.destroy_process_group()


def test_gather_of_shards_equals_full_batch_world2():
    mp.spawn(_worker, args=(2, _free_port(), 8), nprocs=2, join=True)


def _worker_uneven(rank, world, port, total):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(321)
        bon_full = torch.randn(total, 2, 1024, generator=g)
        cor_full = torch.randn(total, 1, 1024, generator=g)
        lo, hi = shard_bounds(total, rank, world)
        for _ in range(3):        # the rotating result buffers must not alias a result still in use
            bon_all, cor_all = gather_outputs(bon_full[lo:hi].clone(), cor_full[lo:hi].clone(), total=total)
            assert torch.equal(bon_all, bon_full) and torch.equal(cor_all, cor_full)
    finally:
        dist.destroy_process_group()


def test_gather_of_uneven_shards_world2():
    """total % world != 0: shards are padded for the collective and trimmed after."""
    mp.spawn(_worker_uneven, args=(2, _free_port(), 7), nprocs=2, join=True)


def test_shard_bounds_cover_everything_once():
    for total in (0, 1, 7, 32, 256):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_bounds(total, r, world)
                assert 0 <= lo <= hi <= total
                seen += list(range(lo, hi))
            assert seen == list(range(total))
    assert shard_bounds(256, 3, 8) == (96, 128)


def _worker_grads(rank, world, port):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        shapes = [(64, 3, 7, 7), (64,), (256, 64, 1, 1), (2048, 1024), (12,), (5, 5)]
        params, expect = [], []
        for i, shp in enumerate(shapes):
            p = torch.nn.Parameter(torch.zeros(shp))
            per_rank = [torch.randn(shp, generator=torch.Generator().manual_seed(100 * r + i)) for r in range(world)]
            if i != 5:                       # the last parameter is frozen on every rank: no gradient anywhere
                p.grad = per_rank[rank].clone()
                expect.append(sum(per_rank) / world)
            else:
                expect.append(None)
            params.append(p)
        calls = average_gradients(params, bucket_bytes=1 << 20)         # 2048*1024*4 B > 1 MiB: several buckets
        assert calls >= 2
        for p, e in zip(params, expect):
            if e is None:
                assert p.grad is None
            else:
                assert torch.allclose(p.grad, e, rtol=0, atol=1e-6)
    finally:
        dist.destroy_process_group()


def test_average_gradients_world2():
    """Data-parallel training step host logic: bucketed gradient all-reduce == mean of the per-rank gradients."""
    mp.spawn(_worker_grads, args=(2, _free_port()), nprocs=2, join=True)
