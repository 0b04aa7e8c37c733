#!/bin/bash
# One GPU-box visit: diagnostics, parity tests, bench lines (both topologies), ncu launch list and
# full captures of the kernels.  Outputs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
for cfg in "ModelWithoutDropoutTdnn B" "ModelWithoutDropout B" "ModelWithoutDropoutTdnn A"; do
  timeout 120 python tools/diag_gpu.py $cfg 2>&1 | tail -12
done | tee gpurun_out/diag.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
if ! grep -q passed gpurun_out/pytest_gpu.log || grep -q failed gpurun_out/pytest_gpu.log; then echo "tests not green: skipping bench/ncu"; exit 1; fi
timeout 600 python bench.py 2>gpurun_out/bench_tdnn.err | tee gpurun_out/bench_tdnn.json
timeout 300 python bench.py --topology ModelWithoutDropout --no-cpu-baseline 2>gpurun_out/bench_dense.err | tee gpurun_out/bench_dense.json
if [ "$1" == "quick" ]; then exit 0; fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 21 -c 28 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tdnn_ -s 15 -c 5 -o gpurun_out/prof_tdnn -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pool_\|pack_im2col -s 6 -c 2 -o gpurun_out/prof_poolpack -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out
