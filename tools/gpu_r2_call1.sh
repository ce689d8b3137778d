#!/bin/bash
# round-2 GPU call 1: parity suite, default bench, full ncu capture (with source) of the search kernel at the bench launch size
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2c1_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2c1_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c1_bench.log 2> gpurun_out/r2c1_bench.err; echo "bench rc $?" >> gpurun_out/r2c1_bench.err
W=1920 H=1088 F=240 REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_kernel -c 1 -f -o gpurun_out/r2_base_240 python tools/prof_run.py > gpurun_out/r2c1_ncu.log 2>&1; echo "ncu rc $?" >> gpurun_out/r2c1_ncu.log
tail -3 gpurun_out/r2c1_pytest.log; tail -2 gpurun_out/r2c1_bench.log; tail -3 gpurun_out/r2c1_ncu.log
