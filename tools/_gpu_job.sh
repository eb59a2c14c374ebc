mkdir -p gpurun_out/r2c
(timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2c/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c/tests.log)
(timeout 200 ./multi-sdr-gps-sim_b200/gpsb200-sol 2999 5 > gpurun_out/r2c/sol.json 2>&1)
(timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c/bench.json 2> gpurun_out/r2c/bench.err)
(timeout 300 python bench.py --steps 5 --warmup 3 --stream-seconds 3600 --no-cpu-baseline > gpurun_out/r2c/bench_3600_1gpu.json 2> gpurun_out/r2c/bench_3600_1gpu.err)
tail -15 gpurun_out/r2c/tests.log; cat gpurun_out/r2c/sol.json; head -c 2500 gpurun_out/r2c/bench.json; tail -3 gpurun_out/r2c/bench.err; head -c 600 gpurun_out/r2c/bench_3600_1gpu.json; tail -2 gpurun_out/r2c/bench_3600_1gpu.err
