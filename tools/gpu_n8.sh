#!/bin/bash
# dev-time N-GPU call (N = number of visible GPUs): the torchrun launch the driver uses for the scaling run, our arm, short
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
df -h /dev/shm | tail -1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 2 --warmup 3 --e2e-steps 1 > gpurun_out/n${N}_bench.log 2> gpurun_out/n${N}_bench.err; echo "bench rc $?"
tail -c 600 gpurun_out/n${N}_bench.log; tail -5 gpurun_out/n${N}_bench.err
