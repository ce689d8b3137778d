#!/bin/bash
# dev-time 2-GPU call: the torchrun launch the driver uses (our arm and the reference arm), short
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --e2e-steps 2 > gpurun_out/n2_bench.log 2> gpurun_out/n2_bench.err; echo "bench rc $?"
tail -c 1500 gpurun_out/n2_bench.log; tail -5 gpurun_out/n2_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/n2_ref.log 2> gpurun_out/n2_ref.err; echo "ref rc $?"
tail -c 700 gpurun_out/n2_ref.log
