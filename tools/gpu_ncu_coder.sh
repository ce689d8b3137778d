#!/bin/bash
# dev-time GPU call: ncu --set full (with source counters) of the slice coder's kernels in one coder pass (FS pictures, default 240)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export FS=${FS:-240} REPS=1
timeout 300 python tools/coder_bench.py > gpurun_out/ncuc_plain.log 2>&1 || { tail -5 gpurun_out/ncuc_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${KERNELS:-cabac_kernel|syntax_kernel}" -c ${COUNT:-3} -f -o gpurun_out/r2_coder python tools/coder_bench.py > gpurun_out/ncuc.log 2>&1; echo "ncu rc $?"
tail -3 gpurun_out/ncuc.log; ls -la gpurun_out/*.ncu-rep
