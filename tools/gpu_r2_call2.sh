#!/bin/bash
# round-2 GPU call 2: parity suite on the packed angular predictor + streaming API, bench (new JSON), quick kernel timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c2_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2c2_pytest.log
tail -25 gpurun_out/r2c2_pytest.log
F=240 timeout 300 python tools/quick_bench.py > gpurun_out/r2c2_quick.log 2>&1; tail -2 gpurun_out/r2c2_quick.log
F=24 timeout 300 python tools/quick_bench.py >> gpurun_out/r2c2_quick.log 2>&1; tail -1 gpurun_out/r2c2_quick.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2c2_bench.log 2> gpurun_out/r2c2_bench.err; echo "bench rc $?" >> gpurun_out/r2c2_bench.err
tail -c 3000 gpurun_out/r2c2_bench.log; tail -5 gpurun_out/r2c2_bench.err
