#!/bin/bash
# the 64-chain leg alone under different pipeline settings: CFGS="name:VAR=val,..."
mkdir -p gpurun_out
{
for c in $CFGS; do
  name=${c%%:*}; envs=${c#*:}; [ "$envs" = "$c" ] && envs=""
  echo "== $name ${envs}"
  env $(echo $envs | tr ',' ' ') timeout 900 python bench.py --workload chains --steps 1 --warmup 1 2>&1 | tail -1 | cut -c1-400
done
} > gpurun_out/exp_chains.log 2>&1
