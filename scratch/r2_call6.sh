#!/bin/bash
mkdir -p gpurun_out
echo "== ncu full: k_render_scatter at S=1e6"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_scatter -s 6 -c 2 -f -o gpurun_out/r02_render python microbench.py --reps 2 --max-scans 90 > gpurun_out/ncu_render.log 2>&1
ls -la gpurun_out/r02_render.ncu-rep
echo "== ncu full: k_gn_persistent"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gn_persistent -s 40 -c 2 -f -o gpurun_out/r02_gn python bench.py --no-cpu-baseline --no-profile --preroll 10 --warmup 5 --steps 3 > gpurun_out/ncu_gn.log 2>&1
ls -la gpurun_out/r02_gn.ncu-rep
echo "== bench default (driver-like)"
timeout 400 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_err.txt | tee gpurun_out/r02_bench_call6.json | cut -c1-600
