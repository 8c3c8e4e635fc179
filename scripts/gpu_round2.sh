#!/bin/bash
mkdir -p gpurun_out
DWB_TEST_TIMEOUT=600 bash scripts/gpu_check.sh tests/test_gpu_model.py
timeout 600 python bench.py --steps 3 --warmup 3 --variant A --no-cpu-baseline > gpurun_out/bench_A.json 2> gpurun_out/bench_A.err; echo "bench A exit=$?"; tail -c 1800 gpurun_out/bench_A.json; tail -4 gpurun_out/bench_A.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 2 -c 1 -o gpurun_out/prof_attn_tc_v2 python scripts/prof_one.py attn 32 20 1500 1500 0 1 > gpurun_out/prof_attn.log 2>&1; echo "attn prof exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 2 -c 1 -o gpurun_out/prof_gemm_n3840_v3 python scripts/prof_one.py gemm 48000 3840 1280 > gpurun_out/prof_gemm2.log 2>&1; echo "gemm prof exit=$?"
