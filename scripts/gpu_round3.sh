#!/bin/bash
mkdir -p gpurun_out
DWB_TEST_TIMEOUT=300 bash scripts/gpu_check.sh tests/test_gpu_kernels.py
DWB_ATTN_VARIANT=1 timeout 200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention_fwd_tcgen05" -p no:cacheprovider > gpurun_out/test_attn_v1.log 2>&1; echo "attn variant1 tests exit=$?"; tail -2 gpurun_out/test_attn_v1.log
timeout 300 python scripts/bench_kernels.py attn > gpurun_out/bench_attn_v2.log 2>&1; echo "--- variant 2"; grep -E "^\{|rror" gpurun_out/bench_attn_v2.log | cut -c1-200
DWB_ATTN_VARIANT=1 timeout 300 python scripts/bench_kernels.py attn > gpurun_out/bench_attn_v1.log 2>&1; echo "--- variant 1"; grep -E "^\{|rror" gpurun_out/bench_attn_v1.log | cut -c1-200
timeout 300 python scripts/bench_kernels.py logmel > gpurun_out/bench_logmel.log 2>&1; grep -E "^\{|rror" gpurun_out/bench_logmel.log | cut -c1-250
cp gpurun_out/bench_kernels.json gpurun_out/bench_logmel.json
timeout 600 python scripts/bench_encoder.py 64 > gpurun_out/bench_encoder.log 2>&1; tail -2 gpurun_out/bench_encoder.log | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -c 600 gpurun_out/bench.json | cut -c1-600; tail -3 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/one_step.py > gpurun_out/one_step.log 2>&1; echo "launch list exit=$?"; wc -l gpurun_out/launches.csv
