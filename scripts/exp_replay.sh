# (historical: MOGP_LAUNCH_AFTER / MOGP_REPLAY were removed when the cached columns became maintained)
run() { name=$1; shift; env "$@" python bench.py --steps 5 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2_exp5_$name.json 2> gpurun_out/r2_exp5_$name.err; }
run replay0_noprune MOGP_PRUNE=0
run replay1_noprune MOGP_PRUNE=0 MOGP_REPLAY=1
python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2_exp5_driverlike.json 2> gpurun_out/r2_exp5_driverlike.err
