#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711"
timeout 150 $TR bench.py --gpus 2 --steps 40 --warmup 10 --striped --comm fused --no-cpu-baseline 2>gpurun_out/striped_fused.err | grep '^{' > gpurun_out/r01_bench_2gpu_striped_fused.json; echo "fused rc=$?"; cut -c1-260 gpurun_out/r01_bench_2gpu_striped_fused.json
timeout 150 $TR bench.py --gpus 2 --steps 40 --warmup 10 --striped --comm callback --no-cpu-baseline 2>gpurun_out/striped_cb.err | grep '^{' > gpurun_out/r01_bench_2gpu_striped_callback.json; echo "callback rc=$?"; cut -c1-260 gpurun_out/r01_bench_2gpu_striped_callback.json
CUDA_VISIBLE_DEVICES=0 timeout 150 python bench.py --steps 4 --warmup 12 --no-cpu-baseline > gpurun_out/r01_bench_steps4_live_table.json 2>/dev/null; echo "steps4 rc=$?"
tail -3 gpurun_out/striped_fused.err gpurun_out/striped_cb.err
