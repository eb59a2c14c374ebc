mkdir -p gpurun_out/r2t
(timeout 1800 python -m pytest tests -m gpu -x -q -s --durations=5 > gpurun_out/r2t/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2t/tests.log)
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2t/smoke.log 2>&1)
(timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2t/bench_driverlike.json 2> gpurun_out/r2t/bench_driverlike.err)
grep -E "single-block|passed|failed|rc=" gpurun_out/r2t/tests.log | tail -5; cat gpurun_out/r2t/smoke.log | tail -2; python -c "
import json
j=json.loads([l for l in open('gpurun_out/r2t/bench_driverlike.json') if l.startswith('{')][-1]); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], j['gpu_launches'], 'cpu' in str(j.get('cpu_baseline')))"
