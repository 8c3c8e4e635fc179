#!/bin/bash
mkdir -p gpurun_out
for v in 1 21; do
DWB_ATTN_POLY=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 2 -c 1 -f -o gpurun_out/r02_prof_attn_v$v \
   python scripts/prof_one.py attn 32 20 1500 1500 0 1 > gpurun_out/r02_prof_attn_v$v.log 2>&1; echo "attn prof v$v exit=$?"
done
ls -la gpurun_out/*.ncu-rep
