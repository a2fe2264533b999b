#!/bin/bash
# runs the diagnostic stages, one process each, logging to gpurun_out/diag.log
mkdir -p gpurun_out
: > gpurun_out/diag.log
for s in "$@"; do
  timeout 300 python tests/diag/gpu_diag.py $s >> gpurun_out/diag.log 2>&1
  echo "stage $s exit $?" >> gpurun_out/diag.log
done
tail -c 6000 gpurun_out/diag.log
