#!/bin/bash
# First GPU call of the next round (1 GPU, ~6 min): is the tree still green, and what do the opt-in variants buy?
#   gpurun --timeout 900 -- ./scratch/round2_first.sh
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== render variants (digest must equal the default's)"
timeout 400 python scratch/variant_parity.py SUMA_B200_RENDER_VARIANT 1 2 2>&1 | tail -4
echo "== Gauss-Newton variant"
timeout 300 python scratch/variant_parity.py SUMA_B200_ICP_VARIANT 1 2>&1 | tail -3
echo "== bench --prefetch (e2e)"; timeout 300 python bench.py --no-cpu-baseline --no-profile --prefetch 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e'])"
for v in "" "SUMA_B200_RENDER_VARIANT=1" "SUMA_B200_RENDER_VARIANT=2" "SUMA_B200_ICP_VARIANT=1"; do
  echo "== bench $v"
  env $v timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d.get('kernels', {})
print(d['value'], d['e2e']['value'], {n: (v['avg_us'], v['share']) for n, v in list(k.items())[:4]})"
done
echo "== microbench S=1e6"
timeout 400 python microbench.py 2>gpurun_out/r02_micro.err | tee gpurun_out/r02_microbench_first.json | cut -c1-1500
echo "== self comm loop-back"
REPS=3 timeout 120 python scratch/self_comm.py 900 2>&1 | tail -3
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1; lscpu | head -30 >> gpurun_out/r02_topo.txt; numactl -H >> gpurun_out/r02_topo.txt 2>&1
