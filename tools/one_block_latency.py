#!/usr/bin/env python3
"""Median latency of one-block calls (the reference's cadence) through gpsb200_synth_blocks with a
host buffer. usage: one_block_latency.py [channels]   (GPSB200_LANES=0 for the lane = channel kernel)"""
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
gps = importlib.import_module("multi-sdr-gps-sim_b200")
nchan = int(sys.argv[1]) if len(sys.argv) > 1 else 12
ch, nav = gps.synthetic_chans(60, nchan, seed=8)
with gps.Context(nchan, 1) as ctx:
    ctx.set_nav_frames(nav)
    cp = None
    out = np.empty(gps.BLOCK_ELEMS, np.int8)
    times = []
    for b in range(60):
        one = ch[b:b + 1].copy()
        if cp is not None:
            one["carr_phase"][0] = cp
        t0 = time.perf_counter()
        _, cp = ctx.synth_blocks(one, 1, out=out)
        times.append(time.perf_counter() - t0)
    name = ctx.synth_kernel_name(nchan)
t = sorted(times[10:])
print(json.dumps({"channels": nchan, "kernel": name, "one_block_call_ms_median": round(t[len(t) // 2] * 1e3, 3),
                  "min": round(t[0] * 1e3, 3), "max": round(t[-1] * 1e3, 3)}))
