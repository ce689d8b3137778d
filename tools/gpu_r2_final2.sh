#!/bin/bash
# round-2 closing GPU call (after the slice-coder rewrite): whole parity suite, default bench, launch list of the bench command,
# ncu --set full capture of the coder kernels at the bench launch size (the search kernel's binary is unchanged: its capture of
# tools/gpu_r2_final.sh stands)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,driver_version --format=csv > gpurun_out/r2g_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2g_pytest.log
tail -3 gpurun_out/r2g_pytest.log
timeout 900 python bench.py > gpurun_out/r2g_bench.log 2> gpurun_out/r2g_bench.err; echo "bench rc $?" >> gpurun_out/r2g_bench.err
tail -c 600 gpurun_out/r2g_bench.log; tail -2 gpurun_out/r2g_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --e2e-steps 1 > gpurun_out/r2g_ncu_bench.log 2>&1; echo "launch list rc $?"
FS=240 REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cabac_kernel|syntax_kernel|nzmap_kernel" -c 4 -f -o gpurun_out/r2_coder_final python tools/coder_bench.py > gpurun_out/r2g_ncu_coder.log 2>&1; echo "ncu coder rc $?"
ls -la gpurun_out | tail -8
