#!/bin/bash
mkdir -p gpurun_out
run() {
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02m_tmp.json 2> gpurun_out/r02m_tmp.err
  python - "$*" <<PY
import json,sys
try:
    d=json.loads([l for l in open("gpurun_out/r02m_tmp.json") if l.startswith("{")][-1])
    print(sys.argv[1], ":", round(d["ms_per_step"],2), "ms/step; exposed", round(d["exposed_comm_ms"]["value"],2), "no-tail", round(d["exposed_comm_ms"]["ms_per_step_without_tail"],2), "parity", d["ddp_parity"]["rel_err"], d["ddp_parity"]["run_to_run_noise"])
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open("gpurun_out/r02m_tmp.err").read()[-800:])
PY
}
run DWB_SYMM_ALLREDUCE=1
run DWB_SYMM_ALLREDUCE=0
run DWB_SYMM_ALLREDUCE=1 DWB_SYMM_CTAS=8
run DWB_SYMM_ALLREDUCE=1 DWB_SYMM_CTAS=64
run DWB_SYMM_ALLREDUCE=1 DWB_TAIL_OVERLAP=0
