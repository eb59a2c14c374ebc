mkdir -p gpurun_out/r2g
(GPSB200_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2g/bench_trace.json 2> gpurun_out/r2g/bench_trace.err)
(timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g/bench.json 2> gpurun_out/r2g/bench.err)
(timeout 900 python -m pytest tests -m gpu -x -q -k "chain or slice or sliced or hand_over or config1_sky12 or 300s or synthetic or randomized" > gpurun_out/r2g/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2g/tests.log)
tail -3 gpurun_out/r2g/tests.log; tail -40 gpurun_out/r2g/bench_trace.err; head -c 1800 gpurun_out/r2g/bench.json
