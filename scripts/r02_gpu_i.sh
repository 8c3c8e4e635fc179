#!/bin/bash
# exposed optimiser tail at N=1 as a function of the tail grid cap
mkdir -p gpurun_out
for g in 0 48 16; do
  DWB_TAIL_GRID=$g timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02i_tail_$g.json 2> gpurun_out/r02i_tail_$g.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02i_tail_$g.json"))
print("tail_grid $g:", round(d["ms_per_step"],2), "ms/step; exposed", d["exposed_comm_ms"]["value"], "no-tail", d["exposed_comm_ms"]["ms_per_step_without_tail"], "clk", d["clocks"]["sm_mhz"])
PY
done
