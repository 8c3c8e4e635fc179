#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention_fwd" -p no:cacheprovider > gpurun_out/test_attn.log 2>&1; echo "attention fwd tests exit=$?"; tail -2 gpurun_out/test_attn.log
timeout 300 python scripts/bench_kernels.py attn > gpurun_out/bench_attn.log 2>&1; grep -E "^\{|rror" gpurun_out/bench_attn.log | cut -c1-230
