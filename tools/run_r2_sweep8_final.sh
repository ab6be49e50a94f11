#!/bin/bash
# cfg 4 with the final kernels: 4 096 of the 8 242 unique architectures, 16 processes on 8 GPUs
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 16 --master-addr 127.0.0.1 --master-port 29761 \
  -m nb_asr_b200.sweep --limit 4096 --out gpurun_out/r2f_sweep_8gpu_p2.json 2> gpurun_out/r2f_sweep_8gpu_p2.err | tail -1 | tee gpurun_out/r2f_sweep_8gpu_p2.summary.json
