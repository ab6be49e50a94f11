#!/bin/bash
# Where does a grouped-conv tile's time go?  Timing experiments on the chain kernel (results are WRONG with NBASR_CHAIN_DBG set):
#   bits 1|2 no dependency protocol, 4 strided tiles, 16 general epilogue, 32 no epilogue arithmetic, 64 no TMA stores,
#   128 no second output, 256 no gate-bit store, 512 epilogue only drains the accumulator  (never bit 1 without bit 2)
# and on the grouped-conv weight gradient: 1024 no gradient atomics, 2048 one MMA per unit.
# Numbers of round 2: profiles/r2_gconv_epilogue_and_chain_study.txt
for dbg in 0 16 35 67 99 515; do
  echo "dbg=$dbg: $(NBASR_CHAIN_DBG=$dbg python tools/one_chain.py)"
done
for dbg in 0 1024 2048 3072; do
  echo "wgrad dbg=$dbg: $(NBASR_CHAIN_DBG=$dbg python tools/one_gconv_wgrad.py | head -1)"
done
