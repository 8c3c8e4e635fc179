#!/bin/bash
# compute-sanitizer over the kernels added in round 2 (decode step, skinny GEMM, timestamp pick, device collator, tensor-core log-mel)
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 700 $SAN --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_logmel.py -m gpu -q -x -p no:cacheprovider \
  -k "decode or skinny or timestamp or collator or reversed_row_walk or tensor_core_dft or golden" > gpurun_out/r02s_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02s_memcheck.log | tail -3
timeout 400 $SAN --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider \
  -k "decode or skinny or timestamp or collator" > gpurun_out/r02s_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r02s_racecheck.log | tail -3
