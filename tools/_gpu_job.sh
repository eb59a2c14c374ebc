mkdir -p gpurun_out/r2d
(timeout 1800 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2d/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2d/tests.log)
(timeout 120 python tools/d2h_peak.py > gpurun_out/r2d/d2h.txt 2>&1)
(timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d/bench.json 2> gpurun_out/r2d/bench.err)
(timeout 300 python bench.py --steps 5 --warmup 3 --stream-seconds 3600 --no-cpu-baseline > gpurun_out/r2d/bench_3600_1gpu.json 2> gpurun_out/r2d/bench_3600_1gpu.err)
tail -12 gpurun_out/r2d/tests.log; cat gpurun_out/r2d/d2h.txt | tail -5; head -c 2000 gpurun_out/r2d/bench.json; tail -3 gpurun_out/r2d/bench.err; head -c 900 gpurun_out/r2d/bench_3600_1gpu.json; tail -2 gpurun_out/r2d/bench_3600_1gpu.err
