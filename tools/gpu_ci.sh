#!/bin/bash
# Runs each GPU test file in its own process under a hard timeout (a hung kernel must not hang the box lease).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv > gpurun_out/nvsmi.txt 2>&1
for f in "$@"; do
  name=$(basename $f .py)
  echo "=== $f" | tee -a gpurun_out/ci.log
  timeout -s KILL 900 python -m pytest $f -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -40 | tee -a gpurun_out/ci_$name.log
  echo "exit: ${PIPESTATUS[0]}" | tee -a gpurun_out/ci.log
done
