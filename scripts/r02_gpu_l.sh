#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --tb=short -k "generate or timestamp or decode" 2>&1 | tail -8
