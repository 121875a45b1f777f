#!/bin/bash
# Round-2 evidence on one B200 (all outputs under gpurun_out/, copied into profiles/ afterwards)
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,driver_version --format=csv > $O/r02_box.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -5 | tee $O/r02_pytest_gpu.txt
echo "== bench (driver command)"; timeout 600 python bench.py --steps 20 --warmup 5 2>$O/bench_err.txt | tee $O/r02_bench_final.json | cut -c1-300
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>>$O/bench_err.txt | tee $O/r02_bench_reference.json | cut -c1-400
echo "== bench default flags"; timeout 600 python bench.py --no-cpu-baseline 2>>$O/bench_err.txt | tee $O/r02_bench_default_flags.json | cut -c1-200
echo "== semantic / ouster lines"
timeout 600 python bench.py --workload hdl64_2048_semantic --steps 20 --warmup 5 --no-cpu-baseline 2>>$O/bench_err.txt | tee $O/r02_bench_semantic.json | cut -c1-200
timeout 600 python bench.py --workload ouster128_4096_geometric --steps 10 --warmup 5 --preroll 20 --no-cpu-baseline 2>>$O/bench_err.txt | tee $O/r02_bench_ouster.json | cut -c1-200
echo "== microbench"; timeout 400 python microbench.py --max-scans 90 --reps 10 2>$O/r02_micro.err | tee $O/r02_microbench_final.json | cut -c1-300
echo "== live table of the ncu command"; timeout 300 python bench.py --steps 4 --warmup 5 --no-cpu-baseline 2>>$O/bench_err.txt > $O/r02_bench_steps4_live_table.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file $O/r02_launches_raw.csv python bench.py --steps 4 --warmup 5 --no-cpu-baseline --no-profile > $O/ncu_list.log 2>&1
wc -l $O/r02_launches_raw.csv
echo "== ncu full set, steady state"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_render_scatter|k_gn_persistent|k_update_surfels|k_index_scatter|k_render_resolve|k_preprocess_tile|k_compact_update|k_gen_compact" -s 660 -c 14 -f -o $O/r02_full python bench.py --steps 4 --warmup 5 --no-cpu-baseline --no-profile > $O/ncu_full.log 2>&1
ls -la $O/r02_full.ncu-rep
cuobjdump -sass semantic_suma_b200/lib/libsuma_b200.so | grep -E "UTMALDG|SYNCS.ARRIVE|SYNCS.PHASECHK" | head -6 > $O/r02_sass_tma.txt; cat $O/r02_sass_tma.txt
