mkdir -p gpurun_out/r2j
nvidia-smi -L | wc -l; nproc; free -g | head -2
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2j/bench_8gpu.json 2> gpurun_out/r2j/bench_8gpu.err)
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2j/bench_4gpu.json 2> gpurun_out/r2j/bench_4gpu.err)
grep -v Warning gpurun_out/r2j/bench_8gpu.err | tail -12 | cut -c1-200; python -c "
import json
for n in (8,4):
    try:
        j=json.load(open('gpurun_out/r2j/bench_%dgpu.json'%n)); print(n, j['value'], j['ms_per_step'], j['kernels']); print(j['e2e'])
    except Exception as e: print(n,'ERR',e)
"
