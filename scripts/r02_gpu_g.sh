#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -p no:cacheprovider --tb=short -k "integration_stub" 2>&1 | grep -E "^E |assert|Error|passed|failed" | head -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02g_bench_n2.json 2> gpurun_out/r02g_bench_n2.err
echo "bench n2 rc=$?"; tail -c 2500 gpurun_out/r02g_bench_n2.json; tail -3 gpurun_out/r02g_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02g_ref_n2.json 2> gpurun_out/r02g_ref_n2.err
echo "ref n2 rc=$?"; cut -c1-300 gpurun_out/r02g_ref_n2.json
