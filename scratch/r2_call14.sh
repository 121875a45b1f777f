#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench14_err.txt | tee gpurun_out/r02_bench_call14.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['clocks'])
print(d['config']['per_rank'])"
tail -3 gpurun_out/bench14_err.txt
