"""CPU, world_size 2 over gloo: the N>1 host logic of the batch path -- contiguous sharding by image, the
max-over-ranks job time, gathering per-image sizes on rank 0 -- with the oracle standing in for the GPU codec."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qoipp_b200 import sharding


def test_shard_range_is_a_partition():
    for n in (1, 2, 7, 8, 37, 8192):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                a, b = sharding.shard_range(n, r, world)
                assert 0 <= a <= b <= n
                seen += list(range(a, b))
                for k in range(a, b):
                    assert sharding.owner_of(k, n, world) == r
            assert seen == list(range(n))
            sizes = [sharding.shard_range(n, r, world) for r in range(world)]
            assert max(b - a for a, b in sizes) - min(b - a for a, b in sizes) <= 1


def _worker(rank, world, port, n_images, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.pyoracle import Oracle
    from qoipp_b200 import synth

    a, b = sharding.shard_range(n_images, rank, world)
    sizes = torch.zeros(n_images, dtype=torch.int64)
    for k in range(a, b):  # each rank encodes only its own images; nothing is exchanged on the data path
        raw = synth.generate("photo", 24, 16, 4, seed=0x51F0 + k)
        sizes[k] = Oracle.encode(raw, 24, 16, 4).size
    dist.reduce(sizes, dst=0, op=dist.ReduceOp.SUM)  # rank 0 collects the per-image sizes (the only cross-rank data)
    job = sharding.job_time_ms(10.0 + 5.0 * rank)
    dist.barrier()
    if rank == 0:
        out_q.put((sizes.tolist(), job))
    dist.destroy_process_group()


def test_two_ranks_shard_a_batch_over_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_images, world = 7, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, q)) for r in range(world)]
    for p in procs:
        p.start()
    sizes, job = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from oracle.pyoracle import Oracle
    from qoipp_b200 import synth

    want = [Oracle.encode(synth.generate("photo", 24, 16, 4, seed=0x51F0 + k), 24, 16, 4).size for k in range(n_images)]
    assert sizes == want
    assert job == 15.0  # max over ranks
