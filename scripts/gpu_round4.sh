#!/bin/bash
mkdir -p gpurun_out
DWB_TEST_TIMEOUT=400 bash scripts/gpu_check.sh tests/test_gpu_kernels.py
timeout 300 python scripts/bench_kernels.py attn_bwd > gpurun_out/bench_attn_bwd.log 2>&1; grep -E "^\{|rror" gpurun_out/bench_attn_bwd.log | cut -c1-250
