#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention" -p no:cacheprovider > gpurun_out/test_attn.log 2>&1; echo "attention tests exit=$?"; tail -2 gpurun_out/test_attn.log
timeout 300 python scripts/bench_kernels.py attn attn_bwd > gpurun_out/bench_attn.log 2>&1; grep -E "^\{|rror" gpurun_out/bench_attn.log | cut -c1-230
timeout 600 python bench.py --steps 3 --warmup 3 --variant A --no-cpu-baseline > gpurun_out/bench_A.json 2> gpurun_out/bench_A.err; echo "bench A exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_A.json')); print({k:d[k] for k in ('value','ms_per_step','ms_per_step_eager')}, d['roofline']['achieved'])"; tail -3 gpurun_out/bench_A.err
