#!/bin/bash
# Round-2 multi-GPU measurements on one 8 x B200 box (run from the repo root or a snapshot of it):
#   1. 2-rank NCCL data-parallel parity test   2. bench.py at N = 8 (cfg 2 value + cfg 3 archs)
#   3. cfg 4: the architecture sweep over all 8 242 unique candidates, sharded over the 8 GPUs
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_6_graph_dp.py -m gpu -q -k two_rank 2>&1 | tail -6 | tee $O/r2_dp2_parity_test.log
for n in 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29730 + n)) \
    bench.py --gpus $n --steps 20 --warmup 5 > $O/r2_bench_dp$n.json 2> $O/r2_bench_dp$n.err
  python - <<PY
import json
d = json.loads(open('$O/r2_bench_dp$n.json').read().strip().splitlines()[-1])
print('N=$n', round(d['value'], 1), 'utt/s', round(d['ms_per_step'], 3), 'ms/step  e2e', round(d['e2e']['value'], 1), 'cfg3', {k: (round(v['ms_per_step'], 2), round(v['value'])) for k, v in (d.get('cfg3') or {}).items()})
PY
done
# cfg 4: every unique architecture once; P processes per GPU
for P in ${SWEEP_PROCS:-1 2}; do
  L=$([ "$P" == "1" ] && echo 0 || echo 2048)          # the second pass (more processes per GPU) only samples 2 048 candidates
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $((8 * P)) --master-addr 127.0.0.1 --master-port $((29750 + P)) \
    -m nb_asr_b200.sweep --limit $L --out $O/r2_sweep_8gpu_p$P.json --out-pickle $O/r2_sweep_8gpu_p$P.pickle 2> $O/r2_sweep_8gpu_p$P.err | tail -1 | tee $O/r2_sweep_8gpu_p$P.summary.json
done
ls -la $O | tail -12
