#!/bin/bash
# regenerates profiles/r02_cusim_runs.txt: the -m gpu suite on the CPU executor (tests/cusim) in its variations
cd "$(dirname "$0")/.."
OUT=profiles/r02_cusim_runs.txt
LIBASAN=$(gcc -print-file-name=libasan.so)
{
echo "# the library's CUDA sources on the CPU executor (tests/cusim, DESIGN.md section 2b); $(nproc) host cores; commit $(git rev-parse --short HEAD)"
echo; echo "## 1. whole -m gpu suite (everything but the two-process test file): python -m pytest tests -q -m gpu --cusim"
python -m pytest tests -q -m gpu --cusim -p no:cacheprovider --durations=8 2>&1 | tail -16
for o in reverse random:1 random:7; do
echo; echo "## 2. lane / warp execution order CUSIM_ORDER=$o (parity + late tests without the two 120-scan loop-closure runs)"
CUSIM_ORDER=$o python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_late.py tests/test_golden.py -q -m gpu --cusim -p no:cacheprovider -k "not loop_closure" 2>&1 | tail -2
done
echo; echo "## 3. AddressSanitizer build of the executor (memcheck of every kernel the suite reaches): CUSIM_ASAN=1 LD_PRELOAD=libasan.so"
CUSIM_ASAN=1 LD_PRELOAD=$LIBASAN ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_late.py tests/test_golden.py tests/test_kitti_io.py tests/test_gpu_cpp_shim.py -q -m gpu --cusim -p no:cacheprovider 2>&1 | tail -2
echo; echo "## 4. B200-sized cooperative grid CUSIM_SMS=148 (296 co-resident blocks of k_gn_persistent, pixel-cache path)"
CUSIM_SMS=148 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_late.py -q -m gpu --cusim -p no:cacheprovider -k "not loop_closure" 2>&1 | tail -2
echo; echo "## 5. in-kernel peer exchange on N ranks (N contexts / host threads), bits of the un-striped run on every rank after every scan"
for n in 2 4 8; do python tests/cusim/multirank_check.py $n 2>&1 | tail -1; done
python tests/cusim/multirank_check.py 4 900 64 5 jump 2>&1 | tail -1
python tests/cusim/multirank_check.py 8 900 64 5 jump 2>&1 | tail -1
echo "# CUSIM_SMS=16: the stripe fits the shared-memory pixel cache (2 ranks: the 4-deep instantiation, 4 ranks: the 2-deep one)"
for n in 2 4; do CUSIM_SMS=16 python tests/cusim/multirank_check.py $n 900 64 5 jump 2>&1 | tail -1; done
echo "# 128x4096 (BASELINE.json configs[3]) striped over 8 ranks, 2 scans"
python tests/cusim/multirank_check.py 8 4096 128 2 2>&1 | tail -1
} > $OUT 2>&1
echo done >> $OUT
