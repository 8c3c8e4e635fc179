#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -p no:cacheprovider --tb=short -k "integration_stub" 2>&1 | grep -E "^E |assert|Error|passed|failed" | head -20
timeout 600 python scripts/debug_det.py 2>&1 | tail -30
