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


def _link_worker(rank, world, port, nblk, q):
    """Hand-over protocol, host side (what bench.py does over NCCL, here over gloo): all-gather of the slices'
    closed-form links -> GUESSED incoming states; then the EXACT states rank to rank (send/recv)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    chans, _ = gps.synthetic_chans(nblk, 12, seed=9)
    chans["prn"][25:, 3] = 0                     # slot 3 sets inside rank 1's slice (world 2: slices of 20)
    chans["prn"][10:, 5] = 31                    # slot 5 takes a new satellite inside rank 0's slice
    chans["carr_phase"][10, 5] = 0.123456789
    lo, hi = gps.sharding.slice_bounds(nblk, world, rank)
    link = gps.slice_link_host(chans[lo:hi])
    flat = np.concatenate([np.asarray(link.prn_first, np.float64), np.asarray(link.prn_last, np.float64),
                           np.asarray(link.reset_inside, np.float64), np.asarray(link.first_phase, np.float64),
                           np.asarray(link.value, np.float64)])
    mine = torch.from_numpy(flat)
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    prn, ph = None, None
    for r in range(rank):
        v = allv[r].numpy().reshape(5, 32)
        lk = gps.SliceLink()
        for c in range(32):
            lk.prn_first[c], lk.prn_last[c], lk.reset_inside[c] = int(v[0, c]), int(v[1, c]), int(v[2, c])
            lk.first_phase[c], lk.value[c] = float(v[3, c]), float(v[4, c])
        prn, ph = gps.link_apply(lk, 12, prn, ph)
    # exact state: received from the previous rank, advanced through the own slice, sent on
    state = torch.zeros(24, dtype=torch.float64)
    if rank > 0:
        dist.recv(state, src=rank - 1)
        eprn, eph = state[:12].numpy().astype(np.int32), state[12:].numpy().copy()
        part = gps.sharding.seed_slice(chans[lo:hi], {"prn": eprn}, eph)
    else:
        eprn, eph, part = None, None, chans[lo:hi]
    end = gps.carrier_chain(part, threads=2)
    if rank + 1 < world:
        dist.send(torch.from_numpy(np.concatenate([part["prn"][-1].astype(np.float64), end])), dst=rank + 1)
    if rank > 0:
        q.put((prn, ph, eprn, eph, end))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_hand_over_guessed_then_exact_state():
    nblk, world = 40, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_link_worker, args=(r, world, port, nblk, q)) for r in range(world)]
    for p in ps:
        p.start()
    gprn, gph, eprn, eph, end = q.get(timeout=120)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    chans, _ = gps.synthetic_chans(nblk, 12, seed=9)
    chans["prn"][25:, 3] = 0
    chans["prn"][10:, 5] = 31
    chans["carr_phase"][10, 5] = 0.123456789
    exact_mid = gps.carrier_chain(chans[:20], threads=1)
    # the guessed incoming state of rank 1: right satellites, phases within 1e-9 cycles of the exact chain
    assert np.array_equal(gprn, chans["prn"][19])
    d = np.abs(gph - exact_mid)
    assert np.all(np.minimum(d, 1.0 - d) < 1e-9), d
    # the exact state that travelled, and the end of the stream
    assert np.array_equal(eph, exact_mid)
    assert np.array_equal(end, gps.carrier_chain(chans, threads=1))
