#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -p no:cacheprovider --tb=short -k "skinny or decode or generate" 2>&1 | tail -5
timeout 600 python scripts/bench_generate.py 2>&1 | tail -4
