#!/bin/bash
# tests + bench + micro-bench in one gpurun call
mkdir -p gpurun_out
bash scripts/gpu_check.sh "$@"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python scripts/bench_kernels.py gemm attn > gpurun_out/bench_kernels.log 2>&1; grep -E "^\{|Error|error" gpurun_out/bench_kernels.log | cut -c1-220
