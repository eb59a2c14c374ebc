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


@pytest.mark.ref
@pytest.mark.parametrize("bits", [8, 16])
def test_reference_sink_code_runs_unmodified_on_our_fifo(bits, tmp_path):
    """Drop-in claim of INTEGRATION.md section 3, compiled and run: the reference's own sdr.c (dispatch table,
    sdr.c:35-87) and sdr_iqfile.c (writer thread, sdr_iqfile.c:22-77) are built unmodified and linked with
    libgpsb200.so INSTEAD OF fifo.o (oracle/Makefile: ref_sinkfeed). A producer replays a stream through
    fifo_acquire / fifo_enqueue exactly as gps_thread_ep does; the sink's iqdata.bin must be that stream --
    all of it (the stock fifo.c would lose buffers 1..6, fifo.c:163-168)."""
    exe = os.path.join(scenario.ROOT, "oracle", "_ref", "ref_sinkfeed")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(scenario.ROOT, "oracle")])
    rng = np.random.default_rng(bits)
    dt = np.int8 if bits == 8 else np.int16
    info = np.iinfo(dt)
    stream = rng.integers(info.min, info.max + 1, size=23 * 600000, dtype=dt)
    src = tmp_path / "in.bin"
    stream.tofile(src)
    out = subprocess.run([exe, str(src), str(bits)], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-400:]
    got = np.fromfile(tmp_path / "iqdata.bin", dtype=dt)
    assert got.size == stream.size and np.array_equal(got, stream)
