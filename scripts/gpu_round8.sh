#!/bin/bash
mkdir -p gpurun_out
DWB_TEST_TIMEOUT=300 bash scripts/gpu_check.sh tests/test_gpu_gemm.py tests/test_gpu_kernels.py tests/test_gpu_logmel.py 2>&1 | grep -E "exit=|passed|failed"
timeout 400 python scripts/bench_kernels.py gemm attn attn_bwd logmel > gpurun_out/bench_all.log 2>&1; grep -E "^\{|rror" gpurun_out/bench_all.log | cut -c1-230
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench B exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','ms_per_step_eager')}, d['roofline']['achieved'])"
