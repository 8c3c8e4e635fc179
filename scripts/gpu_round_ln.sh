#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "layernorm or layer_norm or ln" 2>&1 | tail -3
echo "== sized to the row (64 regs, 4 CTAs/SM)"; timeout 200 python scripts/bench_kernels.py ln 2>&1 | tail -2
echo "== v[16] (old: 2 CTAs/SM)"; DWB_LN_WIDE=1 timeout 200 python scripts/bench_kernels.py ln 2>&1 | tail -2
