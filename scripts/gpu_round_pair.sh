#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu -p no:cacheprovider -x -k "pair" > gpurun_out/test_gemm_pair.log 2>&1; echo "pair tests rc=$?"; tail -8 gpurun_out/test_gemm_pair.log
timeout 600 python scripts/bench_kernels.py gemm 2>&1 | tail -14
