mkdir -p gpurun_out/r2r
(timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2r/bench_1gpu.json 2> gpurun_out/r2r/bench_1gpu.err)
(timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --depth 1 > gpurun_out/r2r/bench_1gpu_d1.json 2> gpurun_out/r2r/bench_1gpu_d1.err)
(BENCH_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e > gpurun_out/r2r/bench_2gpu.json 2> gpurun_out/r2r/bench_2gpu.err)
python -c "
import json
for n in ('1gpu','1gpu_d1','2gpu'):
    try:
        j=json.loads([l for l in open('gpurun_out/r2r/bench_%s.json'%n) if l.startswith('{')][-1]); print(n, j['value'], j['ms_per_step'], j.get('value_kernels_only'), j['kernels']['host_chain_ms_per_step'], j['kernels']['chain_fallback_blocks_per_step']); print('  e2e', j['e2e'].get('value'), j['e2e'].get('ms_per_step'), j['e2e'].get('output_equals_resident_run'))
    except Exception as e: print(n,'ERR',e)
"; tail -3 gpurun_out/r2r/bench_1gpu.err | cut -c1-300; tail -3 gpurun_out/r2r/bench_2gpu.err | cut -c1-300
