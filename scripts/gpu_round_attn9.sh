#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "attention" > gpurun_out/test_attn9.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/test_attn9.log
echo "== BK=64 double-buffered"; timeout 300 python scripts/bench_kernels.py attn 2>&1 | tail -4
echo "== BK=128"; DWB_ATTN_FWD_BK=128 timeout 300 python scripts/bench_kernels.py attn 2>&1 | tail -4
