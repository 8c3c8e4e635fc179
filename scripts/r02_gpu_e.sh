#!/bin/bash
mkdir -p gpurun_out
for v in 20 21; do
  DWB_ATTN_POLY=$v timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --tb=short -k "attention_fwd_tcgen05" 2>&1 | tail -3
done
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -p no:cacheprovider --tb=short -k "integration_stub" 2>&1 | tail -3
timeout 900 python scripts/bench_attn_variants.py 0 1 20 21 22 23 2>&1 | tail -8
