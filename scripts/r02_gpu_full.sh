#!/bin/bash
# round-2 validation: CPU suite, full GPU suite, smoke, default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/r02_full_pytest.log 2>&1
tail -6 gpurun_out/r02_full_pytest.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","exposed_comm_ms")})
print("e2e", d["e2e"]); print("roofline", {k:d["roofline"][k] for k in ("achieved","frac","gemm_share_of_step")})
print("variants", d.get("variants")); print("logmel", d.get("logmel")); print("config5", d.get("config5")); print("gpu_reference", d.get("gpu_reference")); print("cpu", d.get("cpu_baseline"))
PY
