#!/bin/bash
mkdir -p gpurun_out
run() {
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r02o_tmp.json 2> gpurun_out/r02o_tmp.err
  python - "$*" <<PY
import json,sys
d=json.loads([l for l in open("gpurun_out/r02o_tmp.json") if l.startswith("{")][-1])
print(sys.argv[1], ":", round(d["ms_per_step"],2), "ms/step; exposed", round(d["exposed_comm_ms"]["value"],2), "no-tail", round(d["exposed_comm_ms"]["ms_per_step_without_tail"],2), "e2e", round(32/d["e2e"]["value"]*1e3,2), "clk", d["clocks"])
PY
}
run DWB_BENCH_SMI_MS=100
run DWB_BENCH_SMI_MS=0
run DWB_BENCH_SMI_MS=200
run DWB_BENCH_SMI_MS=500
