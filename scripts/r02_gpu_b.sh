#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --tb=short -k "pipelined or state_dict or generate or decode or upstream or out_of_range" > gpurun_out/r02b_pytest.log 2>&1
tail -5 gpurun_out/r02b_pytest.log
timeout 600 python scripts/bench_generate.py > gpurun_out/r02b_generate.log 2>&1; tail -5 gpurun_out/r02b_generate.log
