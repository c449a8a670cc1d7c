#!/usr/bin/env bash
# Runs on the B200 box (via gpurun): GPU parity tests with per-test timeouts, logs into gpurun_out/.
set -uo pipefail
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for t in "$@"; do
  name=$(basename "$t" .py)
  timeout 900 python -m pytest "$t" -q -m gpu --timeout=300 -x -p no:cacheprovider > "gpurun_out/$name.log" 2>&1
  r=$?
  echo "== $t exit $r"; tail -15 "gpurun_out/$name.log"
  [[ $r -ne 0 ]] && rc=$r
done
exit $rc
