#!/bin/bash
mkdir -p gpurun_out
run() {
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r02m_tmp.json 2> gpurun_out/r02m_tmp.err
  python - "$*" <<PY
import json,sys
try:
    d=json.loads([l for l in open("gpurun_out/r02m_tmp.json") if l.startswith("{")][-1])
    x=d["exposed_comm_ms"]
    print(sys.argv[1], ":", round(d["ms_per_step"],2), "ms/step; wait", round(x["value"],3), "tail", round(x["tail_ms_on_side_stream"],2), "E", round(x["graph_E_ms"],2), "D", round(x["graph_D_ms"],2))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open("gpurun_out/r02m_tmp.err").read()[-800:])
PY
}
run DWB_SYMM_ALLREDUCE=1
run DWB_SYMM_ALLREDUCE=0
run DWB_SYMM_ALLREDUCE=1 DWB_SYMM_CTAS=16
run DWB_SYMM_ALLREDUCE=0 DWB_TAIL_OVERLAP=0
