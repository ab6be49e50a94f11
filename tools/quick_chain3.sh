#!/bin/bash
for arch in default c7d2_skips; do
  NBASR_CHAIN_DBG=${DBG:-8} timeout 300 python bench.py --arch $arch --steps 10 --warmup 3 --profile --no-cpu-baseline --no-extra 2>gpurun_out/full_prof_${arch}_dbg${DBG:-8}.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
f=d['roofline']['families']
print('dbg${DBG:-8} $arch', 'step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), {k:(v['ms'],v['n']) for k,v in f.items() if 'gconv' in k})"
  grep "gconv C" gpurun_out/full_prof_${arch}_dbg${DBG:-8}.txt
done
