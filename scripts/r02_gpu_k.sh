#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logmel_dft -s 2 -c 1 -f -o gpurun_out/r02_prof_logmel_dft python scripts/prof_logmel.py > gpurun_out/r02_prof_logmel.log 2>&1; echo "exit=$?"
ls -la gpurun_out/r02_prof_logmel_dft.ncu-rep
