#!/bin/bash
mkdir -p gpurun_out
python scripts/debug_overlap.py 2>&1 | tail -6
run() {
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02i_tmp.json 2> gpurun_out/r02i_tmp.err
  python - "$*" <<PY
import json,sys
d=json.load(open("gpurun_out/r02i_tmp.json"))
print(sys.argv[1], ":", round(d["ms_per_step"],2), "ms/step; exposed", round(d["exposed_comm_ms"]["value"],2), "no-tail", round(d["exposed_comm_ms"]["ms_per_step_without_tail"],2), "clk", d["clocks"]["sm_mhz"])
PY
}
run DWB_TAIL_OVERLAP=0 DWB_TAIL_GRID=0
run DWB_TAIL_OVERLAP=1 DWB_TAIL_GRID=0
run DWB_TAIL_OVERLAP=1 DWB_TAIL_GRID=32
run DWB_TAIL_OVERLAP=1 DWB_TAIL_GRID=96
