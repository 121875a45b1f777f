#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu_call15.txt
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench15_err.txt | tee gpurun_out/r02_bench_call15.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'])
print({n:(v['launches_per_step'],v['avg_us']) for n,v in d['kernels'].items()})"
tail -3 gpurun_out/bench15_err.txt
