#!/bin/bash
# round-2 final GPU call: parity suite, default bench, launch list of the bench command, full ncu capture of the search kernel at the
# bench launch size (240 pictures), per-phase cycle profile is not part of it (needs a -DWB_PROFILE build)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,driver_version --format=csv > gpurun_out/r2f_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2f_pytest.log
tail -3 gpurun_out/r2f_pytest.log
timeout 900 python bench.py > gpurun_out/r2f_bench.log 2> gpurun_out/r2f_bench.err; echo "bench rc $?" >> gpurun_out/r2f_bench.err
tail -c 400 gpurun_out/r2f_bench.log; tail -2 gpurun_out/r2f_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --e2e-steps 1 > gpurun_out/r2f_ncu_bench.log 2>&1; echo "launch list rc $?"
W=1920 H=1088 F=240 REPS=1 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:search_kernel -c 1 -f -o gpurun_out/r2_final_240 python tools/prof_run.py > gpurun_out/r2f_ncu.log 2>&1; echo "ncu rc $?" >> gpurun_out/r2f_ncu.log
tail -3 gpurun_out/r2f_ncu.log
ls -la gpurun_out | tail -12
