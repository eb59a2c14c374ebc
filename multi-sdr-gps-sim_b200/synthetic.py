"""Formulaic synthetic channel parameters (what the reference's 10 Hz host path,
gps.c:2731-2765, would hand to the sample loop) for benchmarks and parity tests.
No ephemerides involved: Doppler, code phase, NAV position and gain follow simple
closed forms with the statistics of a real constellation."""
import numpy as np

from .api import CHAN_DTYPE


def synthetic_chans(nblk, nchan, seed=1, active=None, fmax=5000.0, gain_lo=0.28, gain_hi=1.0, block0=0):
    """-> (chans[nblk, nchan] CHAN_DTYPE, nav[1, nchan, 60] uint32).
    Doppler within +-fmax drifting slowly, f_code tied to f_carr (gps.c:2044), code phase and
    NAV position from a common time base (gps.c:2046-2058), random 30-bit NAV words.
    block0: index of the first block in the scenario (time-slice sharding generates slices of
    one and the same scenario); carr_phase is only meaningful in scenario block 0."""
    rng = np.random.default_rng(seed)
    ch = np.zeros((nblk, nchan), CHAN_DTYPE)
    nav = rng.integers(0, 1 << 30, size=(1, nchan, 60), dtype=np.uint32)
    prns = rng.permutation(32)[:nchan] + 1
    f0 = rng.uniform(-fmax, fmax, nchan)
    fdot = rng.uniform(-0.8, 0.8, nchan)               # Hz per second
    ms0 = rng.uniform(6000.0, 6600.0, nchan)            # ms into the NAV buffer (gps.c:2046)
    g0 = rng.uniform(gain_lo, gain_hi, nchan)
    phase0 = rng.uniform(0, 1, nchan)
    b = np.arange(block0, block0 + nblk, dtype=np.float64)[:, None]
    t = 0.1 * b
    f = f0[None, :] + fdot[None, :] * t
    ms = ms0[None, :] + 1000.0 * np.mod(t, 29.0) * (1.0 + f / 1575.42e6)   # stays inside the 60-word buffer
    ims = np.floor(ms).astype(np.int64)
    ch["prn"] = prns[None, :]
    ch["f_carr"] = f
    ch["f_code"] = 1.023e6 + f * (1.0 / 1540.0)
    ch["code_phase"] = (ms - ims) * 1023.0
    ch["iword"] = ims // 600
    ch["ibit"] = (ims % 600) // 20
    ch["icode"] = ims % 20
    ch["gain"] = g0[None, :] * (1.0 + 0.001 * np.sin(0.01 * b + np.arange(nchan)[None, :]))
    ch["carr_phase"][0] = phase0 if block0 == 0 else 0.0
    if active is not None:
        ch["prn"][:, ~np.asarray(active, dtype=bool)] = 0
    return ch, nav
