#!/bin/bash
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540 + RANDOM % 200)) tools/dp_step_time.py 30 2>&1 | grep "^world"
