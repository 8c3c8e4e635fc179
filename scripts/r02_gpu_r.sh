#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --tb=short -k "reversed_row_walk" 2>&1 | tail -4
for r in 0 1 0 1; do echo row_walk=$r; DWB_ROW_WALK=$r DWB_TAIL_OVERLAP=1 timeout 300 python scripts/debug_phases.py 2>&1 | grep "overlap=1"; done
