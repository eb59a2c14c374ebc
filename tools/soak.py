#!/usr/bin/env python3
"""Long randomized differential run on a B200 (not part of the timed test suite).

  python tools/soak.py [--cases N] [--chains M] [--seed S]

1. N one- to four-block cases (1-2 blocks: the host-resolved small-call path; 3-4: block probes, span chaining and the
   exact run checkpoints on the device) with parameters far outside a real constellation (the generator of
   tests/test_gpu_parity.py::test_randomized_differential_vs_oracle with fresh seeds): CUDA path
   vs the CPU oracle, bit for bit, samples and carried-out carrier phases.
2. M long carrier chains (3000 blocks x 32 channels, random Doppler scale per chain): the two-level
   parallel-in-time device chain (block probes, span chaining, host scan over span summaries) vs the sequential exact
   host walk.
Prints one summary line per part; exit code 1 on the first mismatch."""
import argparse
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
gps = importlib.import_module("multi-sdr-gps-sim_b200")
import scenario  # noqa: E402  (tests/scenario.py: oracle driver)


kernels = {}


def one_case(rng, case, seed):
    nchan = int(rng.choice([1, 3, 8, 12, 16, 20, 32]))
    ss = int(rng.choice([1, 2]))
    nblk = int(rng.choice([1, 2, 3, 4]))
    ch, _ = gps.synthetic_chans(nblk, nchan, seed=seed + case)
    nframes = 3
    nav = rng.integers(0, 1 << 32, size=(nframes, nchan, 60), dtype=np.uint32)
    ch["nav_frame"] = rng.integers(0, nframes, size=(nblk, 1))
    scale = rng.choice([1.0, 6.0, 0.01, 1e-5, 2.5, 0.3])
    ch["f_carr"] *= scale
    ch["f_code"] = 1.023e6 + ch["f_carr"] / 1540.0
    rate = rng.random()
    if rate < 0.2:                  # anywhere in the range k_synth_lanes accepts (its carry-point arithmetic)
        ch["f_code"][:] = rng.uniform(1.01571e6, 1.03019e6, size=nchan)
    elif rate < 0.25:               # outside it: the context falls back to k_synth for good
        ch["f_code"][:, int(rng.integers(0, nchan))] = rng.choice([1.0e6, 1.04e6, 1.0155e6, 1.0303e6])
    ch["gain"] = rng.uniform(0.0, 1.2, size=ch["gain"].shape) * rng.choice([1.0, 0.02])
    edge = rng.integers(0, 4, size=nchan)
    ch["code_phase"][:, edge == 1] = np.nextafter(1023.0, 0)
    ch["code_phase"][:, edge == 2] = 0.0
    ch["icode"][:, edge == 3] = 19
    ch["ibit"][:, edge == 3] = 29
    ch["iword"][:, edge == 3] = rng.integers(0, 59)
    ch["carr_phase"][0] = rng.choice([0.0, np.nextafter(1.0, 0), 0.5, 2.0 ** -40, rng.random()], size=nchan)
    ch["prn"][:, rng.random(nchan) < 0.15] = 0
    want, carr = scenario.oracle_run(ch, nav, ss)
    with gps.Context(nchan, nblk, max_nav_frames=nframes) as ctx:
        ctx.set_nav_frames(nav)
        out, cp = ctx.synth_blocks(ch, ss)
        kernels[ctx.synth_kernel_name(nchan)] = kernels.get(ctx.synth_kernel_name(nchan), 0) + 1
    return np.array_equal(out, want) and np.array_equal(cp, carr), (case, nchan, ss, nblk, float(scale))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=400)
    ap.add_argument("--chains", type=int, default=12)
    ap.add_argument("--seed", type=int, default=777000)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    t0 = time.time()
    for case in range(a.cases):
        ok, what = one_case(rng, case, a.seed)
        if not ok:
            print("MISMATCH in differential case", what)
            return 1
    print("differential: %d cases bit-exact vs the oracle (seed %d) in %.0f s; synthesis kernel of the cases: %s"
          % (a.cases, a.seed, time.time() - t0, kernels))
    t0 = time.time()
    fallbacks = 0
    for k in range(a.chains):
        ch, _ = gps.synthetic_chans(3000, 32, seed=a.seed + 5000 + k)
        ch["f_carr"] *= rng.choice([1.0, 0.1, 3.0, 1e-3, 8.0])
        ch["f_code"] = 1.023e6 + ch["f_carr"] / 1540.0
        with gps.Context(32, 1000) as ctx:
            got = ctx.carrier_chain(ch)
        want = gps.carrier_chain(ch, threads=16)
        if not np.array_equal(got, want):
            print("MISMATCH in carrier chain", k)
            return 1
    print("carrier chains: %d chains x 3000 blocks x 32 channels, device probe + fix-up == sequential host walk, %.0f s"
          % (a.chains, time.time() - t0))
    return 0


if __name__ == "__main__":
    sys.exit(main())
