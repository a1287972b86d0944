#!/bin/bash
# weight-space refit (MOGP_WS_REFIT=1): parity tests, per-cluster trace, bench A/B
mkdir -p gpurun_out
{
export MOGP_WS_REFIT=1
timeout 600 python -m pytest tests/test_engine_parity.py tests/test_headline_parity.py tests/test_chain_parity.py -m gpu -x -q -k "refit or golden or reproducible or notebook or cfg1" 2>&1 | grep -v Warning | tail -12
MOGP_REFIT_TRACE=1 timeout 300 python scripts/refit_lanes.py 32 2>&1 | tail -38
timeout 600 python bench.py --steps 6 --warmup 5 --no-secondary 2>&1 | tail -1
} > gpurun_out/exp_ws.log 2>&1
