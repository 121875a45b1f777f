#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu --tb=short 2>&1 | tail -30 | tee gpurun_out/r02_pytest_gpu_call5.txt
for occ in 4 3 5; do
  echo "== bench RENDER_OCC=$occ"
  SUMA_B200_RENDER_OCC=$occ timeout 300 python bench.py --no-cpu-baseline --preroll 0 --warmup 20 --steps 60 2>gpurun_out/bench_err_$occ.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d.get('kernels', {})
print(d['value'], d['e2e']['value'], d.get('gpu_launches'), {n: (v['avg_us'], v['share']) for n, v in list(k.items())[:14]})"
done
echo "== microbench S=1e6 (OCC 4)"
timeout 400 python microbench.py --max-scans 90 2>gpurun_out/r02_micro.err | tee gpurun_out/r02_microbench_call5.json | cut -c1-900
echo "== bench with preroll (new default)"
timeout 400 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>gpurun_out/bench_err_pre.txt | tee gpurun_out/r02_bench_preroll.json | cut -c1-1500
