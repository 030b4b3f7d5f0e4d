#!/bin/bash
# One full default bench line at N GPUs (what the driver's scaling step runs):  bash scripts/run_scale.sh N [extra bench args]
N=${1:-2}; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"
