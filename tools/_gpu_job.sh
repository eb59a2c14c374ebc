mkdir -p gpurun_out/r2e
(timeout 900 python -m pytest tests -m gpu -x -q -k "chain or slice or sliced or hand_over or config1_sky12 or 300s or cli" > gpurun_out/r2e/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2e/tests.log)
(timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e/bench.json 2> gpurun_out/r2e/bench.err)
(timeout 120 python tools/d2h_peak.py > gpurun_out/r2e/d2h.txt 2>&1)
tail -4 gpurun_out/r2e/tests.log; cat gpurun_out/r2e/d2h.txt | tail -2; head -c 2600 gpurun_out/r2e/bench.json; tail -3 gpurun_out/r2e/bench.err
