#!/usr/bin/env python3
"""Target for `ncu -k regex:k_synth -s 1 -c 1`: two identical device-path calls (the first warms up), 600 blocks.
usage: ncu_one_call.py [channels] [iq16]"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

gps = importlib.import_module("multi-sdr-gps-sim_b200")
nchan = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ss = 2 if len(sys.argv) > 2 and sys.argv[2] == "iq16" else 1
nblk = 600
ch, nav = gps.synthetic_chans(nblk, nchan, seed=7)
out = torch.empty(nblk * gps.BLOCK_ELEMS, dtype=torch.int8 if ss == 1 else torch.int16, device="cuda")
with gps.Context(nchan, nblk) as ctx:
    ctx.set_nav_frames(nav)
    for _ in range(2):
        ctx.synth_blocks_device(ch, ss, out.data_ptr())
        torch.cuda.synchronize()
    print(ctx.synth_kernel_name(nchan))
