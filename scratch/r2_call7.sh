#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu --tb=short 2>&1 | tail -30 | tee gpurun_out/r02_pytest_gpu_call7.txt
for occ in 4 3; do
  echo "== microbench S=1e6 OCC=$occ"
  SUMA_B200_RENDER_OCC=$occ timeout 400 python microbench.py --max-scans 90 --reps 10 2>gpurun_out/r02_micro.err | tee gpurun_out/r02_microbench_call7_occ$occ.json | cut -c1-420
done
echo "== bench (driver-like)"
timeout 400 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>gpurun_out/bench_err.txt | tee gpurun_out/r02_bench_call7.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d.get('kernels', {})
print(d['value'], d['e2e']['value'], d.get('gpu_launches'), {n: (v['avg_us'], v['launches_per_step']) for n, v in list(k.items())[:14]})"
echo "== ncu source of render at S=1e6 (launch 30+)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_scatter -s 75 -c 1 -f -o gpurun_out/r02_render_v4 python microbench.py --reps 2 --max-scans 90 > gpurun_out/ncu_render.log 2>&1
ls -la gpurun_out/r02_render_v4.ncu-rep
