#!/bin/bash
# round-2 GPU call 3: full ncu capture (with source) of the search kernel at the bench launch size (240 pictures), launch list of a bench step
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
W=1920 H=1088 F=240 REPS=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:search_kernel -c 1 -f -o gpurun_out/r2_ang4_240 python tools/prof_run.py > gpurun_out/r2c3_ncu.log 2>&1; echo "ncu rc $?" >> gpurun_out/r2c3_ncu.log
tail -3 gpurun_out/r2c3_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra --e2e-steps 1 > gpurun_out/r2c3_ncu_bench.log 2>&1; echo "launch list rc $?"
ls -la gpurun_out
