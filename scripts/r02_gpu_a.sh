#!/bin/bash
# round 2, first GPU pass: full GPU test suite (no -x: see every failure), then the default bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/r02a_bench.json; tail -5 gpurun_out/r02a_bench.err
