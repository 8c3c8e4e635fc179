#!/bin/bash
mkdir -p gpurun_out
DWB_TEST_TIMEOUT=600 bash scripts/gpu_check.sh tests/test_gpu_gemm.py tests/test_gpu_model.py 2>&1 | grep -v Warning | tail -30
timeout 600 python scripts/bench_kernels.py gemm > gpurun_out/bench_gemm_v5.log 2>&1; grep "48000\|4096" gpurun_out/bench_gemm_v5.log | cut -c1-230
cp gpurun_out/bench_kernels.json gpurun_out/bench_kernels_v5.json
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
