#!/bin/bash
REPS=1 timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,sm__inst_executed.avg.per_cycle_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_cbu.sum,sm__inst_executed_pipe_adu.sum,sm__inst_executed_pipe_uniform.sum,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:gconv_ --csv python tools/one_chain.py 2>&1 | grep -v "^==" > gpurun_out/r2_chain_inst.csv
tail -n +1 gpurun_out/r2_chain_inst.csv | python -c "
import csv,sys,collections
rows=list(csv.reader(sys.stdin))
h=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
H=rows[h]; k=H.index('Kernel Name'); mn=H.index('Metric Name'); mv=H.index('Metric Value'); idc=H.index('ID')
d=collections.OrderedDict()
for r in rows[h+1:]:
    if len(r)<=mv: continue
    d.setdefault((r[idc], r[k].split('(')[0][-22:]), {})[r[mn]]=r[mv]
for kk,v in d.items():
    if 'pack' in kk[1]: continue
    print(kk, {a.replace('sm__inst_executed_pipe_','p_').replace('.sum','').replace('.avg.pct_of_peak_sustained_active','%'):b for a,b in v.items()})
"
