#!/bin/bash
# Run each GPU test file in its own process (a device trap in one file must not poison the others).
# Usage (under gpurun): bash tools/gpu_check.sh [files...]; logs -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
files=("$@")
if [ ${#files[@]} -eq 0 ]; then files=(tests/test_*_gpu.py); fi
rc=0
for f in "${files[@]}"; do
  name=$(basename "$f" .py)
  echo "=== $f"
  timeout 600 python -m pytest "$f" -q -m gpu --tb=short -p no:cacheprovider > "gpurun_out/$name.log" 2>&1
  r=$?
  tail -n 25 "gpurun_out/$name.log"
  echo "=== $f exit $r"
  if [ $r -ne 0 ]; then rc=$r; fi
done
exit $rc
