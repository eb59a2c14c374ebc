"""Shared helpers for the parity tests: golden fixtures -> C-ABI inputs, synthetic
channel-parameter generators, CRC comparison."""
import importlib
import os
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
gps = importlib.import_module("multi-sdr-gps-sim_b200")


def load_golden(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def golden_chans(g, nblocks=None):
    """Golden dump -> (chans[nblk, C] CHAN_DTYPE, nav_frames[nf, C, 60])."""
    ch = g["chans"]
    if nblocks is not None:
        ch = ch[:nblocks]
    out = np.zeros(ch.shape, gps.CHAN_DTYPE)
    for f in ("prn", "iword", "ibit", "icode", "f_carr", "f_code", "carr_phase", "code_phase", "gain"):
        out[f] = ch[f]
    out["nav_frame"] = g["nav_frame_of_block"][:ch.shape[0], None]
    return out, g["nav_frames"]


def crc_blocks(stream):
    nblk = stream.size // gps.BLOCK_ELEMS
    s = stream[:nblk * gps.BLOCK_ELEMS].reshape(nblk, gps.BLOCK_ELEMS)
    return np.array([zlib.crc32(np.ascontiguousarray(s[b]).tobytes()) for b in range(nblk)], np.uint32)


def synthetic_chans(nblk, nchan, seed=1, active=None, fmax=5000.0, gain_lo=0.28, gain_hi=1.0):
    """Formulaic-random channel parameters with the statistics of a real scenario:
    Doppler within +-fmax drifting slowly, f_code tied to f_carr (gps.c:2044), code phase
    and NAV position consistent with a common time base, random 30-bit NAV words."""
    rng = np.random.default_rng(seed)
    ch = np.zeros((nblk, nchan), gps.CHAN_DTYPE)
    nav = rng.integers(0, 1 << 30, size=(1, nchan, 60), dtype=np.uint32)
    prns = rng.permutation(32)[:nchan] + 1
    f0 = rng.uniform(-fmax, fmax, nchan)
    fdot = rng.uniform(-0.8, 0.8, nchan)               # Hz per second
    ms0 = rng.uniform(6000.0, 6600.0, nchan)            # ms into the NAV buffer (gps.c:2046)
    g0 = rng.uniform(gain_lo, gain_hi, nchan)
    for b in range(nblk):
        t = 0.1 * b
        f = f0 + fdot * t
        ms = ms0 + 1000.0 * t * (1.0 + f / 1575.42e6)
        ims = np.floor(ms).astype(np.int64)
        ch["prn"][b] = prns
        ch["f_carr"][b] = f
        ch["f_code"][b] = 1.023e6 + f * (1.0 / 1540.0)
        ch["code_phase"][b] = (ms - ims) * 1023.0
        ch["iword"][b] = ims // 600
        ch["ibit"][b] = (ims % 600) // 20
        ch["icode"][b] = ims % 20
        ch["gain"][b] = g0 * (1.0 + 0.001 * np.sin(0.01 * b + np.arange(nchan)))
        ch["carr_phase"][b] = rng.uniform(0, 1, nchan) if b == 0 else 0.0
    if active is not None:
        ch["prn"][:, ~np.asarray(active, dtype=bool)] = 0
    return ch, nav


def oracle_run(ch, nav, sample_size):
    """Run the CPU oracle over chans (chaining carrier phase like the product API)."""
    import oracle_lib
    nblk, nchan = ch.shape
    outs = []
    carr = np.zeros(nchan)
    prev = np.zeros(nchan, np.int32)
    for b in range(nblk):
        row = ch[b]
        cp = np.where((row["prn"] == prev) & (b > 0), carr, row["carr_phase"])
        chans = oracle_lib.make_chans(row, nav[int(row["nav_frame"][0])], cp)
        iq = oracle_lib.synth_block(chans)
        outs.append(iq if sample_size == 2 else oracle_lib.quantize8(iq))
        carr = np.array([c.carr_phase if c.prn > 0 else 0.0 for c in chans])
        prev = row["prn"].copy()
    return np.concatenate(outs), carr
