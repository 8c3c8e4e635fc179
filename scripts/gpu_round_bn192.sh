#!/bin/bash
mkdir -p gpurun_out
DWB_TEST_TIMEOUT=600 bash scripts/gpu_check.sh tests/test_gpu_gemm.py 2>&1 | tail -5
timeout 600 python scripts/bench_kernels.py gemm > gpurun_out/bench_gemm_v6.log 2>&1; grep "'M': 4096\|'M': 5120\|'M': 1280" gpurun_out/bench_gemm_v6.log | cut -c1-235
cp gpurun_out/bench_kernels.json gpurun_out/bench_kernels_v6.json
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
