#!/bin/bash
# cached columns kept current (MOGP_TMAINT=1, default; two passes in flight) vs one correction per batch + re-scoring (=0)
mkdir -p gpurun_out
{
[ -n "$SKIP_TESTS" ] || timeout 900 python -m pytest tests/test_chain_parity.py -m gpu -x -q -k "lookahead_window or rank_n" 2>&1 | tail -15
for c in ${CFG_LIST:-1:3 1:2 0:2}; do
  tm=${c%%:*}; pd=${c##*:}
  echo "== MOGP_TMAINT=$tm MOGP_PIPELINE=$pd"
  MOGP_TMAINT=$tm MOGP_PIPELINE=$pd timeout 900 python bench.py --steps ${STEPS:-5} --warmup 5 --no-secondary --no-cpu-baseline 2>&1 | tail -1
done
} > gpurun_out/exp_tmaint.log 2>&1
