#!/bin/bash
# round-end evidence on ONE B200: parity tests, ncu captures (launch list + full sets of the top kernels), bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r01_box.txt
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r01_pytest_gpu.txt
B="python bench.py --steps 4 --warmup 12 --no-cpu-baseline --no-profile"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01_launches_final.csv $B > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_render_scatter|k_update_surfels|k_render_resolve|k_gen_compact|k_index_scatter' --launch-skip 100 -c 10 -f -o gpurun_out/r01_top $B > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'k_icp_persistent' --launch-skip 13 -c 2 -f -o gpurun_out/r01_icp $B > gpurun_out/ncu_icp.log 2>&1
echo "icp capture rc=$?"
for r in r01_top r01_icp; do
  [ -f gpurun_out/$r.ncu-rep ] && ncu -i gpurun_out/$r.ncu-rep --page raw --csv --print-units base > gpurun_out/$r.raw.csv 2>/dev/null
done
cat gpurun_out/r01_top.raw.csv > gpurun_out/r01_all.raw.csv 2>/dev/null
[ -f gpurun_out/r01_icp.raw.csv ] && tail -n +3 gpurun_out/r01_icp.raw.csv >> gpurun_out/r01_all.raw.csv
python profiles/ncu_traffic.py gpurun_out/r01_all.raw.csv profiles/dram_traffic.json "ncu --set full --clock-control none, bench.py --steps 4 --warmup 12 (scans 13-16 of the bench sequence), round 1" > gpurun_out/traffic.log 2>&1
cp profiles/dram_traffic.json gpurun_out/dram_traffic.json 2>/dev/null
timeout 400 python bench.py > gpurun_out/r01_bench_final.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-700 gpurun_out/r01_bench_final.json
