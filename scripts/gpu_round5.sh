#!/bin/bash
mkdir -p gpurun_out
DWB_TEST_TIMEOUT=500 bash scripts/gpu_check.sh tests/test_gpu_kernels.py tests/test_gpu_model.py
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench B exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','ms_per_step_eager')}, d['roofline']['achieved'])"
timeout 600 python bench.py --steps 3 --warmup 3 --variant A --no-cpu-baseline > gpurun_out/bench_A.json 2> gpurun_out/bench_A.err; echo "bench A exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_A.json')); print({k:d[k] for k in ('value','ms_per_step','ms_per_step_eager')}, d['roofline']['achieved'])"; tail -3 gpurun_out/bench_A.err
timeout 600 python bench.py --impl hf_gpu --steps 3 --warmup 2 > gpurun_out/bench_hf_gpu.json 2> gpurun_out/bench_hf_gpu.err; echo "hf_gpu exit=$?"; tail -c 600 gpurun_out/bench_hf_gpu.json; tail -3 gpurun_out/bench_hf_gpu.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc_kernel -s 1 -c 1 -o gpurun_out/prof_attn_bwd_tc python scripts/prof_one.py attn_bwd 32 20 1500 1500 0 1 > gpurun_out/prof_attn_bwd.log 2>&1; echo "attn bwd prof exit=$?"
