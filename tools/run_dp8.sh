#!/bin/bash
N=${1:-8}
run() { env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540 + RANDOM % 200)) tools/dp_step_time.py 30 2>&1 | grep "^world"; }
run NBASR_DP_BUCKETS=0
run NBASR_DP_BUCKETS=1
run NBASR_DP_BUCKETS=1 NCCL_MAX_CTAS=8
run NBASR_DP_BUCKETS=1 NCCL_MAX_CTAS=4
run NBASR_DP_BUCKETS=1 NCCL_MAX_CTAS=16
