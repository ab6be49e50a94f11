#!/bin/bash
# compare the persistent stream-K weight-gradient GEMM with the one-tile-per-CTA-pair kernel (NBASR_WGRAD_V1=1)
mkdir -p gpurun_out
unset NBASR_WGRAD_V1
timeout 600 python -m pytest tests/test_gpu_9_sm100.py tests/test_gpu_5_model.py -m gpu -x -q 2>&1 | tail -5
for arch in default linear_skips; do
  timeout 300 python bench.py --arch $arch --steps 10 --warmup 3 --profile --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
f=d['roofline']['families']
print('persist', '$arch', 'step', round(d['ms_per_step'],3), 'gemm_wgrad', f['gemm_wgrad'], 'gemm_tn', f['gemm_tn'])"
done
