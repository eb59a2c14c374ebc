"""Time the host scenario engine (300 s x 32 channels) for a few thread counts."""
import importlib, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gps = importlib.import_module("multi-sdr-gps-sim_b200")
td = tempfile.mkdtemp(); nav = os.path.join(td, "sky32.nav")
subprocess.check_call([sys.executable, os.path.join(ROOT, "oracle", "gen_rinex.py"), "--nsat", "32", "--out", nav])
for thr in ("1", "4", "16"):
    os.environ["GPSB200_SCENARIO_THREADS"] = thr
    best = 1e9
    for _ in range(8):
        t = time.time(); gps.scenario(nav, 35.681298, 139.766247, 10.0, 300, max_chan=32, start=(2024, 1, 7, 2, 0, 0.0)); best = min(best, time.time() - t)
    print("scenario engine, 300 s x 32 ch, %s thread(s): %.1f ms" % (thr, best * 1e3))
