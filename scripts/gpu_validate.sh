#!/bin/bash
# Round-end validation on one B200: every GPU test file (own process each), smoke(), bench variant B (default) and A.
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -m pytest tests -x -q -m "not gpu" -p no:cacheprovider > gpurun_out/test_cpu.log 2>&1; echo "cpu tests rc=$?"; tail -2 gpurun_out/test_cpu.log
DWB_TEST_TIMEOUT=900 bash scripts/gpu_check.sh tests/test_gpu_gemm.py tests/test_gpu_kernels.py tests/test_gpu_logmel.py tests/test_gpu_model.py tests/test_gpu_fullsize.py > gpurun_out/gpu_check.log 2>&1
cat gpurun_out/summary.txt
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 1200 python bench.py > gpurun_out/bench_B.json 2> gpurun_out/bench_B.err; echo "bench B rc=$?"; cut -c1-140 gpurun_out/bench_B.json
timeout 1200 python bench.py --variant A --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_A.json 2> gpurun_out/bench_A.err; echo "bench A rc=$?"; cut -c1-140 gpurun_out/bench_A.json
