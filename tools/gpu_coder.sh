#!/bin/bash
# dev-time GPU call for the slice-coder work: parity (slice_data byte-identical vs the oracle) of the default build and of the
# fallback build, then the coder stage's time for every wrenc_b200/lib*.so variant (tools/coder_bench.py)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${TAG:-coder}
SEL='golden or cif_two or qp_sweep or random_and_flat or single_ctu or streaming or extreme or second_walk or arena or decoded_gpu or pinned'
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_reference_clips.py -m gpu -x -q -k "$SEL or clip" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest default rc $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
WRENC_B200_LIB=$GRAFT_REPO_ROOT/wrenc_b200/lib_old.so timeout 600 python -m pytest tests/test_gpu_search.py -m gpu -x -q -k "golden or qp_sweep or extreme" > gpurun_out/${TAG}_pytest_old.log 2>&1; echo "pytest lib_old rc $?" >> gpurun_out/${TAG}_pytest_old.log
tail -3 gpurun_out/${TAG}_pytest_old.log
rm -f gpurun_out/${TAG}_bench.log
for lib in wrenc_b200/libwrenc_b200.so wrenc_b200/lib_*.so; do
  [ -f "$lib" ] || continue
  WRENC_B200_LIB=$GRAFT_REPO_ROOT/$lib timeout 300 python tools/coder_bench.py >> gpurun_out/${TAG}_bench.log 2>&1
done
cat gpurun_out/${TAG}_bench.log
