#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_logmel.py -q -m gpu -p no:cacheprovider > gpurun_out/test_logmel.log 2>&1; echo "logmel tests exit=$?"; tail -3 gpurun_out/test_logmel.log
timeout 300 python scripts/bench_kernels.py logmel > gpurun_out/bench_logmel.log 2>&1; grep -E "^\{|rror" gpurun_out/bench_logmel.log | cut -c1-250
timeout 300 ncu --set full --clock-control none --import-source on -k regex:logmel_kernel -s 1 -c 1 -o gpurun_out/prof_logmel python scripts/prof_one.py logmel 256 > gpurun_out/prof_logmel.log 2>&1; echo "logmel prof exit=$?"
