#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711"
timeout 240 $TR bench.py --gpus 2 --steps 40 --warmup 10 --striped --comm fused 2>/dev/null | grep '^{' > gpurun_out/r01_bench_2gpu_striped_fused.json; cut -c1-400 gpurun_out/r01_bench_2gpu_striped_fused.json
timeout 240 $TR bench.py --gpus 2 --steps 40 --warmup 10 --striped --comm callback 2>/dev/null | grep '^{' > gpurun_out/r01_bench_2gpu_striped_callback.json; cut -c1-400 gpurun_out/r01_bench_2gpu_striped_callback.json
timeout 240 $TR bench.py --gpus 2 --steps 40 --warmup 10 2>/dev/null | grep '^{' > gpurun_out/r01_bench_2gpu_weak.json; cut -c1-400 gpurun_out/r01_bench_2gpu_weak.json
