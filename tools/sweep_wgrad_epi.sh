#!/bin/bash
# effect of the split-K cost model's epilogue constant on the dense weight-gradient GEMMs (bench.py --profile families)
for e in 1.0 3.0 6.0 12.0; do
  NBASR_WGRAD_EPI_US=$e python bench.py --steps 10 --warmup 3 --profile --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('epi_us=$e', 'step', round(d['ms_per_step'],3), 'gemm_wgrad', d['roofline']['families']['gemm_wgrad'])"
done
