#!/bin/bash
# bench.py over the device work unit (run length in samples); prints one line per setting
for r in ${@:-800 2400 4000 12000}; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --run-samples $r 2>/dev/null |
    R=$r python -c 'import json,sys,os; d=json.loads(sys.stdin.read()); print(os.environ["R"], d["value"], d["kernels"], d["e2e"]["value"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])'
done
