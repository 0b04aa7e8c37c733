#!/bin/bash
# first bring-up on the B200 box: diagnostics for every kernel-addressing variant, then the tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for cfg in "ModelWithoutDropout B 0 1" "ModelWithoutDropoutTdnn B 0 1" "ModelWithoutDropout B 1 1" "ModelWithoutDropout B 1 0" "ModelWithoutDropoutTdnn B 1 1"; do
  timeout 180 python tools/diag_gpu.py $cfg 2>&1 | tail -12
done | tee gpurun_out/diag.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
