#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu --tb=short 2>&1 | tail -12 | tee gpurun_out/r02_pytest_gpu_call11.txt
for cfg in "1 1" "0 0" "1 0" "0 1"; do
  set -- $cfg
  echo "== bench GN_LL=$1 GN_CACHE=$2"
  SUMA_B200_GN_LL=$1 SUMA_B200_GN_CACHE=$2 timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>gpurun_out/bench_err.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d.get('kernels', {})
print(d['value'], d['e2e']['value'], d.get('gpu_launches'), {n: (v['avg_us'], v['launches_per_step']) for n, v in list(k.items())[:5]})"
done
