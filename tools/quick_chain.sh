#!/bin/bash
# chain kernel: parity tests, then step time with and without chaining (NBASR_GCONV_NO_CHAIN=1)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_8_gconv_chain.py -m gpu -x -q 2>&1 | tail -15
for nc in 0 1; do
  for arch in ${ARCHS:-default c7d2_skips}; do
    NBASR_GCONV_NO_CHAIN=$nc timeout 300 python bench.py --arch $arch --steps 10 --warmup 3 --profile --no-cpu-baseline --no-extra 2>gpurun_out/chain_prof_${arch}_$nc.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
f=d['roofline']['families']
print('no_chain=$nc', '$arch', 'step', round(d['ms_per_step'],3), 'launches', d['gpu_launches'], 'gconv', f['gconv'], 'gconv_wgrad', f['gconv_wgrad'], 'floor', (d['roofline'].get('mma_issue_floor') or {}).get('frac_of_measured'))"
  done
done
