#!/bin/bash
# dev-time GPU call: parity suite of the default build, then resident-search throughput of the default build and of every wrenc_b200/lib_*.so variant
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${TAG:-ab}
if [ -z "$NOTEST" ]; then
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
fi
rm -f gpurun_out/${TAG}_var.log
for lib in wrenc_b200/libwrenc_b200.so wrenc_b200/lib_*.so; do
  [ -f "$lib" ] || continue
  for f in ${FS:-240 24}; do
    WRENC_B200_LIB=$GRAFT_REPO_ROOT/$lib F=$f GSTEP=4 timeout 200 python tools/quick_bench.py >> gpurun_out/${TAG}_var.log 2>&1
  done
done
cat gpurun_out/${TAG}_var.log
