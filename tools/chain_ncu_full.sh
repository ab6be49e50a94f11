#!/bin/bash
mkdir -p gpurun_out
python tools/one_chain.py
NBASR_CHAIN_DBG=8 python tools/one_chain.py
REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gconv -c 10 -o gpurun_out/r2_chain_vs_old -f python tools/one_chain.py > gpurun_out/r2_chain_ncu.log 2>&1
tail -3 gpurun_out/r2_chain_ncu.log
