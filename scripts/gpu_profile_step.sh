#!/bin/bash
# ncu evidence for the final build: (1) --set full of the CTA-pair GEMM at the fused-QKV shape, (2) launch list of one KD step
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:gemm_bf16_2cta -s 1 -c 1 -f -o gpurun_out/final_gemm_pair python scripts/prof_one.py gemm 48000 3840 1280 > gpurun_out/final_prof_gemm_pair.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_final.csv python scripts/one_step.py > gpurun_out/one_step_final.log 2>&1
ls -la gpurun_out/final_gemm_pair.ncu-rep gpurun_out/launches_final.csv; tail -2 gpurun_out/one_step_final.log
