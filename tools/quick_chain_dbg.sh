#!/bin/bash
# timing-only experiments on the chain kernel (NBASR_CHAIN_DBG switches off parts of the dependency protocol; never 1 alone)
for dbg in ${DBGS:-0 3 7 8}; do
  NBASR_CHAIN_DBG=$dbg timeout 120 python bench.py --arch default --steps 5 --warmup 3 --profile --no-cpu-baseline --no-extra 2>gpurun_out/chain_dbg_$dbg.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
f=d['roofline']['families']
print('dbg=$dbg', 'step', round(d['ms_per_step'],3), 'gconv', f['gconv'])"
  grep "gconv C" gpurun_out/chain_dbg_$dbg.txt
done
NBASR_GCONV_NO_CHAIN=1 timeout 120 python bench.py --arch default --steps 5 --warmup 3 --profile --no-cpu-baseline --no-extra 2>gpurun_out/chain_dbg_nochain.txt | tail -c 300
grep "gconv C" gpurun_out/chain_dbg_nochain.txt
