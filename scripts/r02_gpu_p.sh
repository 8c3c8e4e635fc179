#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches.csv python scripts/one_step.py > gpurun_out/r02_one_step.log 2>&1
echo "launch list exit=$?"; wc -l gpurun_out/r02_launches.csv; tail -2 gpurun_out/r02_one_step.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:logmel_dft -s 2 -c 1 -f -o gpurun_out/r02_prof_logmel_dft_final python scripts/prof_logmel.py > gpurun_out/r02_prof_logmel_final.log 2>&1; echo "logmel prof exit=$?"
