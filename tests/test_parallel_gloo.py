"""world_size-2 `gloo` test (CPU) of the batch-sharding / output-gather logic in
nnaudio_b200/parallel.py — the N>1 host path.  The transform is injected (the
CPU oracle), because the product kernels are CUDA-only."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import build, run_oracle

from nnaudio_b200.parallel import BatchShardedTransform, shard_bounds


def test_shard_bounds_cover_batch_exactly():
    for total in (0, 1, 5, 64, 1023):
        for world in (1, 2, 3, 8):
            b = shard_bounds(total, world)
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mod = build("STFT", dict(n_fft=256, hop_length=64, output_format="Magnitude"))

        def transform(x):
            return torch.from_numpy(run_oracle("STFT", mod, x.numpy(), {}, dtype=np.float32))

        x_global = torch.from_numpy(
            np.random.RandomState(7).standard_normal((total, 2048)).astype(np.float32))
        sharded = BatchShardedTransform(transform)
        y = sharded(sharded.local_slice(x_global), total=total)
        full = transform(x_global)
        ret[rank] = bool(y.shape == full.shape and torch.equal(y, full))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [4, 5])
def test_two_rank_gather_matches_single_process(total):
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), total, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def _async_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mod = build("STFT", dict(n_fft=256, hop_length=64, output_format="Magnitude"))

        def transform(x):
            return torch.from_numpy(run_oracle("STFT", mod, x.numpy(), {}, dtype=np.float32))

        sharded = BatchShardedTransform(transform)
        rng = np.random.RandomState(11)
        batches = [torch.from_numpy(rng.standard_normal((4, 2048)).astype(np.float32)) for _ in range(3)]
        # the bench.py pattern: the gather of step i is waited for only after step i+1 was issued,
        # with two rotating gather buffers
        ok, pending = True, None
        for i, xg in enumerate(batches):
            work, buf = sharded.forward_async(sharded.local_slice(xg), slot=i & 1)
            if pending is not None:
                pw, pbuf, pfull = pending
                pw.wait()
                ok = ok and torch.equal(pbuf, pfull)
            pending = (work, buf, transform(xg))
        pending[0].wait()
        ok = ok and torch.equal(pending[1], pending[2])
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_two_rank_pipelined_gather_with_rotating_buffers():
    """forward_async (what bench.py runs for N > 1): results of step i stay intact in their slot
    while step i+1 is transformed and gathered into the other slot."""
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_async_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
