#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_6_graph_dp.py -m gpu -q -k two_rank 2>&1 | tail -8 | tee $O/r2f_dp2_parity_test.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29732 \
  bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2f_bench_dp2.json 2> $O/r2f_bench_dp2.err
python - <<PY
import json
d = json.loads(open('$O/r2f_bench_dp2.json').read().strip().splitlines()[-1])
print('N=2', round(d['value'], 1), 'utt/s', round(d['ms_per_step'], 3), 'ms/step  e2e', round(d['e2e']['value'], 1), 'launches', d['gpu_launches'], 'cfg3', {k: (round(v['ms_per_step'], 2), round(v['value'])) for k, v in (d.get('cfg3') or {}).items()})
PY
tail -3 $O/r2f_bench_dp2.err
