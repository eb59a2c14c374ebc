mkdir -p gpurun_out/r2prof
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2prof/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --depth 1 > gpurun_out/r2prof/launches_bench.log 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_synth -c 2 -o gpurun_out/r2prof/prof_synth python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --blocks 600 --depth 1 > gpurun_out/r2prof/prof_synth.log 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_probe|k_chain|k_checkpoints" -c 6 -o gpurun_out/r2prof/prof_walks python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --blocks 600 --depth 1 > gpurun_out/r2prof/prof_walks.log 2>&1)
(timeout 900 python tools/soak.py --cases 300 --chains 8 > gpurun_out/r2prof/soak.txt 2>&1)
ls -la gpurun_out/r2prof; tail -3 gpurun_out/r2prof/soak.txt; tail -2 gpurun_out/r2prof/prof_synth.log
