#!/bin/bash
# line coverage of semantic_suma_b200/csrc under the -m gpu suite on the CPU executor (gcov build of tests/cusim)
cd "$(dirname "$0")/.."
rm -rf tests/cusim/_build/cov
CUSIM_COV=1 python tests/cusim/build_sim.py >/dev/null
CUSIM_COV=1 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_late.py tests/test_golden.py tests/test_kitti_io.py -q -m gpu --cusim -p no:cacheprovider -k "not loop_closure" | tail -1
cd tests/cusim/_build/cov && gcov -b -o . gen/sb_map.cpp gen/sb_icp.cpp gen/sb_preprocess.cpp gen/sb_api.cpp 2>/dev/null | grep -A3 "File '.*semantic_suma_b200/csrc/sb_.*\.cu'" | grep -v "^--" | paste - - - - | sed "s#File '.*/csrc/##"
