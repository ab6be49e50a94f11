#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/bench_gemm.py 2>&1 | tail -12
for arch in default c7d2_skips linear_skips; do
  timeout 300 python bench.py --arch $arch --steps 10 --warmup 3 --profile --no-cpu-baseline --no-extra 2>gpurun_out/full_prof_$arch.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
f=d['roofline']['families']
print('$arch', 'step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), {k:(v['ms'],v['n'],v['tflops']) for k,v in f.items() if k in ('gemm_tn','gemm_wgrad','gconv','gconv_wgrad')})"
done
