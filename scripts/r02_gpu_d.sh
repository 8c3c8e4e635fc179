#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -p no:cacheprovider --tb=short -k "collator or integration_stub" 2>&1 | tail -5
timeout 900 python scripts/bench_attn_variants.py 0 1 2 3 4 10 11 12 13 14 2>&1 | tail -14
