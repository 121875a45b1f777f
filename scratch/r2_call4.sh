#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -60 | tee gpurun_out/r02_pytest_gpu_call4.txt
for occ in 3 2 4; do
  echo "== bench RENDER_OCC=$occ"
  SUMA_B200_RENDER_OCC=$occ timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d.get('kernels', {})
print(d['value'], d['e2e']['value'], d.get('gpu_launches'), {n: (v['avg_us'], v['share']) for n, v in list(k.items())[:14]})"
done
echo "== microbench S=1e6 (OCC 3)"
timeout 400 python microbench.py --max-scans 90 2>gpurun_out/r02_micro.err | tee gpurun_out/r02_microbench_call4.json | cut -c1-1800
