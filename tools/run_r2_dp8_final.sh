#!/bin/bash
# final build on one 8 x B200 box: bench.py --gpus 8 (cfg 2 value + cfg 3 archs at N = 8)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29738 \
  bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2f_bench_dp8.json 2> $O/r2f_bench_dp8.err
python - <<PY
import json
d = json.loads(open('$O/r2f_bench_dp8.json').read().strip().splitlines()[-1])
print('N=8', round(d['value'], 1), 'utt/s', round(d['ms_per_step'], 3), 'ms/step  e2e', round(d['e2e']['value'], 1), 'cfg3', {k: (round(v['ms_per_step'], 2), round(v['value'])) for k, v in (d.get('cfg3') or {}).items()})
PY
