#!/bin/bash
# One GPU-box visit for the committed evidence of the shipped kernels (outputs -> gpurun_out/, tag $1):
#   the default bench line, the ncu launch list of `bench.py --steps 2 --warmup 1`, and `ncu --set full` of ONE extraction
#   step of both topologies (tools/profile_step.py; 3 warm-up steps skipped).
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
tail -c 300 gpurun_out/${TAG}_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-train --no-product --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
for TOPO in ModelWithoutDropoutTdnn ModelWithoutDropout; do
  # per step: 1 torch fill (the L2 flush) + 7 library launches; capture the 4th step
  timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 24 -c 8 -o gpurun_out/${TAG}_prof_${TOPO} -f \
      python tools/profile_step.py $TOPO > gpurun_out/${TAG}_ncu_full_${TOPO}.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu_full_${TOPO}.log
done
ls -la gpurun_out | tail -8
