#!/bin/bash
# where does a grouped-conv tile's time go?  (timing experiments: results are wrong with NBASR_CHAIN_DBG set)
for dbg in 0 96 99 515; do
  echo "dbg=$dbg: $(NBASR_CHAIN_DBG=$dbg python tools/one_chain.py)"
done
