#!/bin/bash
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --profile-from-start off -k regex:"attn_decode|gemm_skinny" -c 6 -f -o gpurun_out/r02_prof_decode python scripts/prof_decode.py > gpurun_out/r02_prof_decode_full.log 2>&1; echo "exit=$?"
ls -la gpurun_out/r02_prof_decode.ncu-rep
