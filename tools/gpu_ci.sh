#!/bin/bash
# Runs each GPU test file in its own process under a hard timeout (a hung kernel must not hang the box lease).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv > gpurun_out/nvsmi.txt 2>&1
for f in "$@"; do
  name=$(basename $f .py)
  echo "=== $f" | tee -a gpurun_out/ci.log
  timeout -s KILL 900 python -m pytest $f -q -m gpu --no-header -p no:cacheprovider > gpurun_out/ci_$name.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/ci_$name.log | tail -30
  
done
