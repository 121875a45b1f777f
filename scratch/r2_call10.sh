#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -40 | tee gpurun_out/r02_pytest_gpu_call10.txt
