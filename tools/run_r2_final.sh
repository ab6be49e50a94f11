#!/bin/bash
# Round-2 final measurements on ONE B200 (outputs under gpurun_out/, summaries copied to profiles/ by hand):
#   1. ncu launch list of two eager steps   2. ncu --set full of one eager step   3. default bench line (with CPU arm)
#   4. --impl reference arm   5. cfg 5 (32 x 3000 frames)
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file $O/r2f_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-extra > $O/r2f_launches.log 2>&1
python tools/ncu_summary.py list $O/r2f_launches.csv $O/r2f_ncu_launches_summary.txt | head -30
timeout 900 ncu --set full --clock-control none -c 700 -o $O/r2f_step_full -f \
  python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-extra > $O/r2f_step_full.log 2>&1
python tools/ncu_summary.py full $O/r2f_step_full.ncu-rep $O/r2_ncu_step_final.json > /dev/null 2>&1
cp $O/r2_ncu_step_final.json profiles/ 2>/dev/null
timeout 900 python bench.py --profile > $O/r2f_bench.json 2> $O/r2f_bench_breakdown.txt
tail -c 600 $O/r2f_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2f_bench_reference_arm.json 2> $O/r2f_bench_reference_arm.err
tail -c 400 $O/r2f_bench_reference_arm.json
timeout 300 python bench.py --batch 32 --frames 3000 --steps 5 --warmup 3 --no-cpu-baseline --no-extra --profile > $O/r2f_bench_long.json 2> $O/r2f_bench_long_breakdown.txt
python -c "
import json
d=json.loads(open('$O/r2f_bench_long.json').read().strip().splitlines()[-1])
print('cfg5 32x3000:', round(d['ms_per_step'],2), 'ms/step', round(d['value'],1), 'utt/s', {k:v['ms'] for k,v in d['roofline']['families'].items() if 'lstm' in k})"
rm -f $O/r2f_step_full.ncu-rep
