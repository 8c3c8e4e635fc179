#!/bin/bash
for d in 0 1 6; do DWB_LOGMEL_DBG=$d timeout 120 python scripts/debug_logmel_time.py 2>&1 | tail -1; done
timeout 300 python -m pytest tests/test_gpu_logmel.py -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -3
