timeout 900 python -m pytest tests/test_gpu_5_model.py tests/test_gpu_6_graph_dp.py tests/test_gpu_8_gconv_chain.py -m gpu -x -q 2>&1 | tail -3
for arch in default c7d2_skips; do
timeout 300 python bench.py --arch $arch --steps 10 --warmup 3 --profile --no-cpu-baseline --no-extra 2>gpurun_out/skipg_prof_$arch.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
f=d['roofline']['families']
print('$arch step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), {k:(v['ms'],v['gbs']) for k,v in f.items() if k in ('gconv','ln_bwd','ln_fwd')}, 'frac', round(d['roofline']['frac'],3))"
done
