#!/bin/bash
# compute-sanitizer (memcheck, racecheck) over the feature front-end kernels alone.
mkdir -p gpurun_out
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/frontend_sanitize_target.py > gpurun_out/sanitize_frontend_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/sanitize_frontend_memcheck.log
tail -4 gpurun_out/sanitize_frontend_memcheck.log
timeout 200 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/frontend_sanitize_target.py > gpurun_out/sanitize_frontend_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/sanitize_frontend_racecheck.log
tail -4 gpurun_out/sanitize_frontend_racecheck.log
