mkdir -p gpurun_out/r2h
(GPSB200_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2h/bench_trace.json 2> gpurun_out/r2h/bench_trace.err)
(timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h/bench.json 2> gpurun_out/r2h/bench.err)
(GPSB200_CHECK_STRIDE=1 timeout 900 python -m pytest tests -m gpu -x -q -k "chain or slice or sliced or hand_over or config1_sky12 or 300s or synthetic or randomized or 3600" > gpurun_out/r2h/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2h/tests.log)
tail -3 gpurun_out/r2h/tests.log; tail -12 gpurun_out/r2h/bench_trace.err; python -c "
import json
j=json.load(open('gpurun_out/r2h/bench.json')); print(j['value'], j['ms_per_step'], j['kernels']); print(j['e2e'])"
