#!/bin/bash
# One GPU-box visit for the feature front end: parity tests, bench line, ncu launch list and full captures.
mkdir -p gpurun_out
timeout 600 python -m pytest ${GPU_TESTS:-tests} -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
if ! grep -q passed gpurun_out/pytest_gpu.log || grep -q failed gpurun_out/pytest_gpu.log; then echo "tests not green: skipping bench/ncu"; exit 1; fi
timeout 100 python tools/frontend_profile.py 2>&1 | tail -3 | tee gpurun_out/frontend_profile.log
[ -n "$SKIP_BENCH" ] || timeout 600 python bench.py 2>gpurun_out/bench_tdnn.err | tee gpurun_out/bench_tdnn.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'cmvn_select|vad_tile' --csv \
    --log-file gpurun_out/frontend_launches.csv python tools/frontend_profile.py 5 > gpurun_out/ncu_frontend_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'cmvn_select|vad_tile' -s 6 -c 4 -o gpurun_out/prof_frontend -f \
    python tools/frontend_profile.py 5 > gpurun_out/ncu_frontend_full.log 2>&1
ls -la gpurun_out | tail -8
