#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_8_gconv_chain.py -m gpu -x -q 2>&1 | tail -5
python tools/one_chain.py
NBASR_CHAIN_DBG=16 python tools/one_chain.py
bash tools/quick_chain5.sh | tail -2 | cut -c1-600
