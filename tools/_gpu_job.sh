mkdir -p gpurun_out/r2i
nvidia-smi -L | head -3
(timeout 600 python -m pytest tests -m gpu -x -q -k "cli" > gpurun_out/r2i/tests_cli.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2i/tests_cli.log)
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2i/bench_2gpu.json 2> gpurun_out/r2i/bench_2gpu.err)
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --impl reference > gpurun_out/r2i/bench_2gpu_ref.json 2> gpurun_out/r2i/bench_2gpu_ref.err)
tail -3 gpurun_out/r2i/tests_cli.log; tail -25 gpurun_out/r2i/bench_2gpu.err | cut -c1-300; head -c 3000 gpurun_out/r2i/bench_2gpu.json; head -c 400 gpurun_out/r2i/bench_2gpu_ref.json
