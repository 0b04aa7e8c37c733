#!/bin/bash
# ncu evidence for the training step: launch list of two steps + full captures of the tensor-core kernels.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 170 --csv --log-file gpurun_out/train_launches.csv \
    python tools/bench_train.py ModelWithoutDropoutTdnn 64 400 5000 2 > gpurun_out/ncu_train_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_pair -s 10 -c 5 -o gpurun_out/prof_wgrad -f \
    python tools/bench_train.py ModelWithoutDropoutTdnn 64 400 5000 2 > gpurun_out/ncu_train_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tdnn_pair -s 18 -c 9 -o gpurun_out/prof_train_pair -f \
    python tools/bench_train.py ModelWithoutDropoutTdnn 64 400 5000 2 > gpurun_out/ncu_train_full2.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:blk_col_sums\|bn_relu_bwd\|bn_apply\|adam -s 20 -c 8 -o gpurun_out/prof_train_hbm -f \
    python tools/bench_train.py ModelWithoutDropoutTdnn 64 400 5000 2 > gpurun_out/ncu_train_full3.log 2>&1
ls -la gpurun_out | tail -12
