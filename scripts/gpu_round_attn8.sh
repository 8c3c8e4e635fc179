#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "attention" -x > gpurun_out/test_attn8.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/test_attn8.log
echo "== 8 softmax warps"; timeout 300 python scripts/bench_kernels.py attn 2>&1 | tail -12
echo "== 4 softmax warps"; DWB_ATTN_FWD_WARPS=4 timeout 300 python scripts/bench_kernels.py attn 2>&1 | tail -12
