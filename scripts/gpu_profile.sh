#!/bin/bash
mkdir -p gpurun_out
# launch list of one bench step (skip the 3 warm-up steps + model build launches is approximate: capture the tail)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 20000 -c 7000 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list exit=$?"; wc -l gpurun_out/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 2 -c 1 -o gpurun_out/prof_gemm_n3840 \
   python scripts/prof_one.py gemm 48000 3840 1280 > gpurun_out/prof_gemm2.log 2>&1; echo "gemm prof exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 2 -c 1 -o gpurun_out/prof_attn_tc \
   python scripts/prof_one.py attn 32 20 1500 1500 0 1 > gpurun_out/prof_attn.log 2>&1; echo "attn prof exit=$?"
