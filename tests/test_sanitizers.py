"""Race / memory checks: the FIFO under ThreadSanitizer (CPU), the kernels under compute-sanitizer (GPU)."""
import os
import shutil
import subprocess
import sys

import pytest

import scenario

ROOT = scenario.ROOT


def test_fifo_is_race_free_under_tsan(tmp_path):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    exe = tmp_path / "fifo_tsan"
    cmd = [gxx, "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-I", os.path.join(ROOT, "include"),
           "-I/usr/local/cuda/include", "-o", str(exe), os.path.join(ROOT, "tests", "native", "fifo_tsan.cpp"),
           os.path.join(ROOT, "multi-sdr-gps-sim_b200", "csrc", "fifo.cpp"),
           "-L/usr/local/cuda/lib64", "-lcudart", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("TSan build not available here: " + r.stderr[-200:])
    env = dict(os.environ, LD_LIBRARY_PATH="/usr/local/cuda/lib64:" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([str(exe)], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ThreadSanitizer" not in r.stderr


@pytest.mark.gpu
def test_kernels_clean_under_compute_sanitizer():
    cs = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(cs):
        pytest.skip("compute-sanitizer not installed")
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import numpy as np, scenario; "
            "gps = scenario.gps; ch, nav = gps.synthetic_chans(2, 32, seed=5); ctx = gps.Context(32, 2); "
            "ctx.set_nav_frames(nav); out, cp = ctx.synth_blocks(ch, 1); "
            "ch12, nav12 = gps.synthetic_chans(1, 12, seed=6); c2 = gps.Context(12, 1); c2.set_nav_frames(nav12); "
            "o2, _ = c2.synth_blocks(ch12, 2); print('ok', int(np.abs(out).sum()), int(np.abs(o2.astype(np.int64)).sum()))"
            % (ROOT, os.path.join(ROOT, "tests")))
    for tool in ("memcheck", "racecheck"):
        r = subprocess.run([cs, "--tool", tool, "--error-exitcode", "9", sys.executable, "-c", code],
                           capture_output=True, text=True, timeout=1500)
        assert r.returncode == 0, (tool, r.stdout[-1500:], r.stderr[-500:])
        assert "ok" in r.stdout
