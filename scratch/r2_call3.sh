#!/bin/bash
mkdir -p gpurun_out
echo "== tests, default (TMA off at 64x900/2048, on at 128x4096)"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
echo "== tests with TMA forced on"
SUMA_B200_PREP_TMA=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "preprocess or pipeline" 2>&1 | tail -8
for v in 0 1; do
  echo "== bench PREP_TMA=$v"
  SUMA_B200_PREP_TMA=$v timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d.get('kernels', {})
print(d['value'], d['e2e']['value'], {n: (v['avg_us'], v['share']) for n, v in list(k.items())[:12]})"
  echo "== bench ouster PREP_TMA=$v"
  SUMA_B200_PREP_TMA=$v timeout 300 python bench.py --no-cpu-baseline --workload ouster128_4096_geometric --steps 12 --warmup 6 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d.get('kernels', {})
print(d['value'], d['e2e']['value'], {n: (v['avg_us'], v['share']) for n, v in list(k.items())[:12]})"
done
cuobjdump -sass semantic_suma_b200/lib/libsuma_b200.so | grep -c UTMALDG
