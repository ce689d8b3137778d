#!/bin/bash
# dev-time GPU call: parity suite (stops at the first failure), then resident-search throughput at 240 and 24 pictures
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${TAG:-q}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
F=240 timeout 300 python tools/quick_bench.py > gpurun_out/${TAG}_quick.log 2>&1; tail -2 gpurun_out/${TAG}_quick.log
F=24 timeout 300 python tools/quick_bench.py >> gpurun_out/${TAG}_quick.log 2>&1; tail -1 gpurun_out/${TAG}_quick.log
