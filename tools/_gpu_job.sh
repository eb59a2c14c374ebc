mkdir -p gpurun_out/r2n
(timeout 1800 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2n/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2n/tests.log)
(timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2n/bench.json 2> gpurun_out/r2n/bench.err)
tail -9 gpurun_out/r2n/tests.log; python -c "
import json
j=json.loads([l for l in open('gpurun_out/r2n/bench.json') if l.startswith('{')][-1]); print(j['value'], j['ms_per_step'], j.get('value_kernels_only'), j['kernels']); print(j['e2e'])"
