"""Time-slice sharding of one continuous IQ stream across ranks.

A 0.1 s block depends only on its own channel parameters and on each channel's carrier
phase at the block start; everything else is re-seeded per block (gps.c:2046-2060). So
rank r can synthesize blocks [lo_r, hi_r) independently once it knows the exact carrier
phases at block lo_r, which the exact fast-forward provides without synthesizing anything.
No data-path collective is needed; an optional NCCL all-gather collects the finished int8
slices for a single sink."""
import numpy as np

from . import api


def slice_bounds(nblocks_total, world_size, rank):
    """Contiguous, balanced slices: ranks < nblocks % world get one extra block."""
    base, extra = divmod(nblocks_total, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def start_phases(chans_prefix, carr_phase0=None, threads=16, ctx=None):
    """Exact carrier phase of every channel slot after the blocks in chans_prefix[nblk, nchan]
    (same chaining rule as gpsb200_synth_blocks). Host-only: O(#binade crossings) per
    channel-block, multi-threaded over channels inside libgpsb200 (gpsb200_carrier_chain)."""
    if chans_prefix.shape[0] == 0:
        return np.zeros(chans_prefix.shape[1]) if carr_phase0 is None else np.array(carr_phase0, dtype=np.float64)
    if ctx is not None:          # parallel-in-time on the rank's own GPU: milliseconds instead of seconds
        return ctx.carrier_chain(chans_prefix, carr_phase0)
    return api.carrier_chain(chans_prefix, carr_phase0, threads)


def seed_slice(chans_slice, prefix_last_row, phases):
    """Seed the first block of a rank's slice: a slot CONTINUES the carrier phase chained through the
    preceding blocks (`phases`, from start_phases) only where it still holds the satellite it held in the
    block before the slice (`prefix_last_row`, the chans row of block lo-1); a slot that is (re)allocated
    exactly at the slice edge keeps the allocation phase the scenario put into carr_phase
    (allocateChannel, gps.c:2203-2210) -- the same rule gpsb200_synth_blocks applies between the blocks of
    one call. Returns a copy of chans_slice."""
    out = np.array(chans_slice, copy=True)
    if prefix_last_row is None or out.shape[0] == 0:
        return out
    keep = (out["prn"][0] > 0) & (out["prn"][0] == prefix_last_row["prn"])
    out["carr_phase"][0] = np.where(keep, phases, out["carr_phase"][0])
    return out
