#!/bin/bash
# cached columns kept current (MOGP_TMAINT=1, default) vs the one-correction-per-batch + re-scoring path (=0)
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_chain_parity.py -m gpu -x -q -k "lookahead_window or rank_n" 2>&1 | tail -15
for c in ${TM_LIST:-1 0}; do
  echo "== MOGP_TMAINT=$c"
  MOGP_TMAINT=$c timeout 900 python bench.py --steps ${STEPS:-5} --warmup 5 --no-secondary --no-cpu-baseline 2>&1 | tail -1
done
} > gpurun_out/exp_tmaint.log 2>&1
