#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8
nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/r02_topo_8gpu.txt
for N in ${NS:-8 4}; do
  echo "== bench --gpus $N"
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench${N}_err.txt | tee gpurun_out/r02_bench_${N}gpu.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'])
s = d.get('striped') or {}
print({k: (v if not isinstance(v, dict) else {a: v[a] for a in ('ms_per_scan', 'gn_launch_us_avg', 'rows')}) for k, v in s.items() if k in ('solo', 'striped', 'speedup_vs_solo', 'poses_bit_identical_across_ranks_and_modes')})
print([(r['rank'], r['step_ms_min_med_max'], r['clocks']['sm_mhz']) for r in d['config']['per_rank']])"
  tail -3 gpurun_out/bench${N}_err.txt
done
