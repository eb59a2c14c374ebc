mkdir -p gpurun_out/r2p
(BENCH_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 6 --warmup 3 --no-e2e > gpurun_out/r2p/bench_2gpu.json 2> gpurun_out/r2p/bench_2gpu.err)
grep "step phases" gpurun_out/r2p/bench_2gpu.err | cut -c1-330; python -c "
import json
for n in (2,):
    try:
        j=json.loads([l for l in open('gpurun_out/r2p/bench_%dgpu.json'%n) if l.startswith('{')][-1]); print(n, j['value'], j['ms_per_step'], j['kernels']['host_chain_ms_per_step'], j['kernels']['chain_fallback_blocks_per_step'], j['kernels']['k_synth_ms'])
    except Exception as e: print(n,'ERR',e)
"; tail -3 gpurun_out/r2p/bench_2gpu.err | cut -c1-300
