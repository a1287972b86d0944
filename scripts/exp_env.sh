#!/bin/bash
# bench A/B over environment settings: CFGS="name:VAR=val,VAR2=val ..." (name "base" = no settings)
mkdir -p gpurun_out
{
for c in $CFGS; do
  name=${c%%:*}; envs=${c#*:}; [ "$envs" = "$c" ] && envs=""
  echo "== $name ${envs}"
  env $(echo $envs | tr ',' ' ') timeout 900 python bench.py --steps ${STEPS:-8} --warmup 5 --no-secondary --no-cpu-baseline 2>&1 | tail -1
done
} > gpurun_out/exp_env.log 2>&1
