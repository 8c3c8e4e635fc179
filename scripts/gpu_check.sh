#!/bin/bash
# Run on the GPU box (via gpurun): every GPU test file in its own process so that one trapping kernel does not
# poison the rest; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for t in "$@"; do
  name=$(basename "$t" .py)
  timeout ${DWB_TEST_TIMEOUT:-420} python -m pytest "$t" -q -m gpu -p no:cacheprovider --durations=6 > "gpurun_out/${name}.log" 2>&1
  code=$?
  echo "$name exit=$code" | tee -a gpurun_out/summary.txt
  tail -n 25 "gpurun_out/${name}.log"
  [ $code -ne 0 ] && rc=1
done
exit $rc
