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


synthetic_chans = importlib.import_module("multi-sdr-gps-sim_b200.synthetic").synthetic_chans


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
