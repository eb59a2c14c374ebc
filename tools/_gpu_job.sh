mkdir -p gpurun_out/r2m
nproc; cat /sys/fs/cgroup/cpu.max; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null | head -2
(GPSB200_TRACE=1 BENCH_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 2 --warmup 1 --no-e2e > gpurun_out/r2m/bench_4gpu.json 2> gpurun_out/r2m/bench_4gpu.err)
grep "gpsb200 dev\|step phases" gpurun_out/r2m/bench_4gpu.err | tail -44 | cut -c1-300
