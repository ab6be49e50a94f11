#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_1_kernels.py tests/test_gpu_5_model.py -m gpu -x -q -k "lstm or model or step or logits" 2>&1 | tail -4
timeout 300 python bench.py --arch default --steps 10 --warmup 3 --profile --no-cpu-baseline --no-extra 2>gpurun_out/lstm_prof.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
f=d['roofline']['families']
print('default step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), {k:(v['ms'],v['n']) for k,v in f.items()})"
