#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713"
A="--workload ouster128_4096_geometric --steps 30 --warmup 10 --no-cpu-baseline --no-profile"
timeout 110 $TR bench.py --gpus 2 $A --striped --comm fused 2>/dev/null | grep '^{' > gpurun_out/r01_bench_2gpu_ouster_striped_fused.json; echo "fused rc=$?"; cut -c1-200 gpurun_out/r01_bench_2gpu_ouster_striped_fused.json
CUDA_VISIBLE_DEVICES=0 timeout 110 python bench.py $A 2>/dev/null > gpurun_out/r01_bench_1gpu_ouster.json; echo "single rc=$?"; cut -c1-200 gpurun_out/r01_bench_1gpu_ouster.json
