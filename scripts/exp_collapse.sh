#!/bin/bash
# A/B of the collapsed-shadow refit (MOGP_COLLAPSE=0 = evaluations on the full cluster)
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_engine_parity.py tests/test_chain_parity.py tests/test_headline_parity.py tests/test_reference_sampler.py -m gpu -x -q 2>&1 | tail -5
for c in 0 1; do
  echo "== MOGP_COLLAPSE=$c"
  MOGP_COLLAPSE=$c timeout 600 python scripts/refit_lanes.py 32 32 2>&1 | tail -3
  MOGP_COLLAPSE=$c timeout 900 python bench.py --steps 6 --warmup 5 --no-secondary --no-cpu-baseline 2>&1 | tail -1
done
} > gpurun_out/exp_collapse.log 2>&1
