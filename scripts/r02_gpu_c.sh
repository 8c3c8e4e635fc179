#!/bin/bash
# full GPU suite, then compute-sanitizer passes over the hand-rolled mbarrier / TMEM / shared-memory protocols (bounded)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/r02c_pytest.log 2>&1
tail -8 gpurun_out/r02c_pytest.log
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $SAN --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gemm.py -m gpu -q -x -p no:cacheprovider \
  -k "decode or collator or kd_loss or layernorm or adamw or embed or tcgen05 or pair or attention or gemm" > gpurun_out/r02c_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02c_memcheck.log | tail -4
timeout 600 $SAN --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider \
  -k "decode or collator or kd_loss or layernorm or attention_fwd_tcgen05" > gpurun_out/r02c_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r02c_racecheck.log | tail -4
