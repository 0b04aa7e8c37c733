#!/bin/bash
# ncu launch list + full captures (source-level) of the layer kernel and the HBM-side kernels.
mkdir -p gpurun_out
TOPO=${1:-ModelWithoutDropoutTdnn}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 27 -c 27 --csv --log-file gpurun_out/launches.csv \
    python bench.py --topology $TOPO --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tdnn_ -s 18 -c 6 -o gpurun_out/prof_tdnn -f \
    python bench.py --topology $TOPO --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pool_\|pack_im2col\|embed_ -s 9 -c 3 -o gpurun_out/prof_poolpack -f \
    python bench.py --topology $TOPO --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out
