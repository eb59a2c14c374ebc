"""The reference program exactly as shipped (own fifo.c / sdr_iqfile.c) against the golden
enqueue stream: its iqdata.bin is the stream with blocks 1..6 missing (fifo.c:163-168).
Needs the prebuilt oracle/_ref/ref_stock12 (built in the container that has /root/reference)."""
import os
import subprocess
import sys
import zlib

import numpy as np
import pytest

import scenario

EXE = os.path.join(scenario.ROOT, "oracle", "_ref", "ref_stock12")


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/ref_stock12 not built")
def test_stock_iqfile_is_enqueue_stream_minus_blocks_1_to_6(tmp_path):
    nav = tmp_path / "sky12.nav"
    subprocess.check_call([sys.executable, os.path.join(scenario.ROOT, "oracle", "gen_rinex.py"),
                           "--nsat", "12", "--out", str(nav)])
    subprocess.check_call([EXE, "-e", str(nav), "-l", "35.681298,139.766247,10.0", "-d", "10"], cwd=tmp_path,
                          stderr=subprocess.DEVNULL)
    s = np.fromfile(tmp_path / "iqdata.bin", dtype=np.int8)
    g = scenario.load_golden("sky12_static_10s_i8")
    keep = [0] + list(range(7, 99))
    assert s.size == len(keep) * 600000
    blocks = s.reshape(len(keep), 600000)
    for row, b in zip(blocks, keep):
        assert zlib.crc32(row.tobytes()) == g["crcs"][b, 0], b
