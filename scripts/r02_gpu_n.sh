#!/bin/bash
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/test_symm_allreduce.py 2>&1 | grep -v "OMP_NUM\|\*\*\*\*\|UserWarning\|return func" | tail -20
