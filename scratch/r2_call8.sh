#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu --tb=short 2>&1 | tail -30 | tee gpurun_out/r02_pytest_gpu_call8.txt
echo "== microbench S=1e6"
timeout 400 python microbench.py --max-scans 90 --reps 10 2>gpurun_out/r02_micro.err | tee gpurun_out/r02_microbench_call8.json | cut -c1-420
echo "== bench (driver-like)"
timeout 400 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>gpurun_out/bench_err.txt | tee gpurun_out/r02_bench_call8.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d.get('kernels', {})
print(d['value'], d['e2e']['value'], d.get('gpu_launches'), {n: (v['avg_us'], v['launches_per_step']) for n, v in list(k.items())[:14]})"
