#!/bin/bash
# ncu evidence for the bench line: (1) launch list of a steady-state sweep (timed region only, first N launches),
# (2) one --set full capture of the dominant kernel in the same region.  Numbers printed by bench.py under ncu are not bench values.
mkdir -p gpurun_out
export MOGP_PROFILE_RANGE=1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c ${NLAUNCH:-1400} --csv \
  --log-file gpurun_out/r2b_launches_steady.csv python bench.py --steps 1 --warmup 6 --no-secondary --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "ncu_rc=$?" >> gpurun_out/ncu_list.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_score_wide -c 4 \
  -o gpurun_out/r2b_score_wide -f python bench.py --steps 1 --warmup 6 --no-secondary --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu_rc=$?" >> gpurun_out/ncu_full.log
ncu -i gpurun_out/r2b_score_wide.ncu-rep --page raw --csv > gpurun_out/r2b_score_wide_raw.csv 2>/dev/null
