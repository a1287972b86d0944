#!/bin/bash
mkdir -p gpurun_out
{
for c in ${CFG_LIST:-32:3 32:2 8:3}; do
  cn=${c%%:*}; pd=${c##*:}
  echo "== CUDA_DEVICE_MAX_CONNECTIONS=$cn MOGP_PIPELINE=$pd"
  CUDA_DEVICE_MAX_CONNECTIONS=$cn MOGP_PIPELINE=$pd timeout 900 python bench.py --steps ${STEPS:-5} --warmup 5 --no-secondary --no-cpu-baseline 2>&1 | tail -1
done
} > gpurun_out/exp_conn.log 2>&1
