#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -q -m gpu --tb=short -k "multi or fused or stripe or loopback or comm or minimize" 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu_2gpu_c.txt
echo "== bench --gpus 2 (replicas + striped record)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench2_err.txt | tee gpurun_out/r02_bench_2gpu_c.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'])
print(json.dumps(d.get('striped'), indent=1))
print([ (r['rank'], r['step_ms_min_med_max'], r['clocks']['sm_mhz'], r['numa_cpus']) for r in d['config']['per_rank']])"
tail -5 gpurun_out/bench2_err.txt
