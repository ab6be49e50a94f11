#!/bin/bash
# 2-GPU A/B of the bucketed (overlapped) gradient exchange against the single all-reduce
for b in 0 1; do
  NBASR_DP_BUCKETS=$b python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29520 + b)) \
    bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/s51_dp2_buckets$b.json 2> gpurun_out/s51_dp2_buckets$b.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/s51_dp2_buckets$b.json').read().strip().splitlines()[-1])
print('buckets=$b', d['n_gpus'], 'GPUs', round(d['value'], 1), 'utt/s', round(d['ms_per_step'], 3), 'ms/step  e2e', round(d['e2e']['value'], 1))
PY
done
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/s51_dp1.json 2> gpurun_out/s51_dp1.err
python -c "
import json
d = json.loads(open('gpurun_out/s51_dp1.json').read().strip().splitlines()[-1])
print('1 GPU', round(d['value'], 1), 'utt/s', round(d['ms_per_step'], 3), 'ms/step')"
