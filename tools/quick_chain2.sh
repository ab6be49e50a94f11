#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_8_gconv_chain.py -m gpu -x -q 2>&1 | tail -3
python tools/one_chain.py
C=1000 OPS=conv7d2,conv7d2,conv7d2 BWD=1 python tools/one_chain.py
DBGS="0" bash tools/quick_chain_dbg.sh 2>&1 | grep -v '^"' | cut -c1-200
