"""Multi-rank host logic on CPU (gloo, world_size 2): time-slice bounds and the exact
slice-start carrier phases that make N ranks produce one continuous stream."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import scenario
from scenario import gps


def test_slice_bounds_partition_the_stream():
    for total in (1, 7, 2999, 35999):
        for world in (1, 2, 4, 8):
            edges = [gps.sharding.slice_bounds(total, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            for (a, b), (c, d) in zip(edges, edges[1:]):
                assert b == c and b - a >= d - c >= 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nblk, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    chans, _ = gps.synthetic_chans(nblk, 12, seed=9)
    lo, hi = gps.sharding.slice_bounds(nblk, world, rank)
    start = gps.sharding.start_phases(chans[:lo], threads=2) if lo > 0 else chans["carr_phase"][0].copy()
    end = gps.carrier_chain(chans[lo:hi], phase_in=start, threads=2)
    t = torch.from_numpy(np.stack([start, end]))
    got = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(got, t)
    if rank == 0:
        q.put([g.numpy() for g in got])
    dist.destroy_process_group()


def test_two_ranks_chain_into_one_continuous_stream():
    nblk, world = 40, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, nblk, q)) for r in range(world)]
    for p in ps:
        p.start()
    got = q.get(timeout=120)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    # rank 1 starts exactly where rank 0 ends, and the last rank ends where a single-process chain ends
    assert np.array_equal(got[0][1], got[1][0])
    chans, _ = gps.synthetic_chans(nblk, 12, seed=9)
    assert np.array_equal(got[1][1], gps.carrier_chain(chans, threads=1))
    assert np.array_equal(got[0][0], chans["carr_phase"][0])
