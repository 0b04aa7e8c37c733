#!/bin/bash
# One GPU-box visit: parity tests, bench lines (both topologies, both tap-addressing modes),
# ncu launch list and one full capture of the fused TDNN layer kernel.  Outputs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py 2>gpurun_out/bench_tdnn.err | tee gpurun_out/bench_tdnn.json
timeout 300 python bench.py --topology ModelWithoutDropout --no-cpu-baseline 2>gpurun_out/bench_dense.err | tee gpurun_out/bench_dense.json
timeout 300 python bench.py --reuse-taps 1 --no-cpu-baseline 2>gpurun_out/bench_tdnn_reuse.err | tee gpurun_out/bench_tdnn_reuse.json
timeout 300 python bench.py --topology ModelWithoutDropout --reuse-taps 1 --no-cpu-baseline 2>gpurun_out/bench_dense_reuse.err | tee gpurun_out/bench_dense_reuse.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 21 -c 28 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tdnn_layer -s 15 -c 5 -o gpurun_out/prof_tdnn -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pool_embed\|pack_im2col -s 6 -c 2 -o gpurun_out/prof_poolpack -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out
