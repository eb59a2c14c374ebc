mkdir -p gpurun_out/r2b
(timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2b/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2b/tests.log)
(timeout 200 ./multi-sdr-gps-sim_b200/gpsb200-sol 2999 5 > gpurun_out/r2b/sol.json 2>&1)
(timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b/bench.json 2> gpurun_out/r2b/bench.err)
tail -15 gpurun_out/r2b/tests.log; cat gpurun_out/r2b/sol.json; head -c 1500 gpurun_out/r2b/bench.json; tail -3 gpurun_out/r2b/bench.err; nproc; free -g | head -2
