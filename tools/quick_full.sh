#!/bin/bash
# whole GPU suite + default / cfg3 step times at the current state
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for arch in default c7d2_skips linear_skips; do
  timeout 300 python bench.py --arch $arch --steps 10 --warmup 3 --profile --no-cpu-baseline --no-extra 2>gpurun_out/full_prof_$arch.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
f=d['roofline']['families']
print('$arch', 'step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'launches/step', d['gpu_launches']/(d['steps']+d['warmup']) if 0 else d['gpu_launches'], {k:(v['ms'],v['n']) for k,v in f.items()})"
done
