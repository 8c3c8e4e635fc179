#!/bin/bash
# full-size parity test + ncu --set full captures of the final attention fwd / bwd and log-mel kernels
mkdir -p gpurun_out
DWB_TEST_TIMEOUT=800 bash scripts/gpu_check.sh tests/test_gpu_fullsize.py
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:attn_fwd_tc_kernel -s 1 -c 1 -f -o gpurun_out/final_attn_fwd python scripts/prof_one.py attn 32 20 1500 1500 0 1 > gpurun_out/final_prof_attn_fwd.log 2>&1
timeout 300 $NCU -k regex:attn_bwd_tc_kernel -s 1 -c 1 -f -o gpurun_out/final_attn_bwd python scripts/prof_one.py attn_bwd 32 20 1500 1500 0 1 > gpurun_out/final_prof_attn_bwd.log 2>&1
timeout 300 $NCU -k regex:logmel -s 1 -c 1 -f -o gpurun_out/final_logmel python scripts/prof_one.py logmel 256 > gpurun_out/final_prof_logmel.log 2>&1
ls -la gpurun_out/*.ncu-rep
