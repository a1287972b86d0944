run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2_exp9_$name.json 2> gpurun_out/r2_exp9_$name.err; }
run prehi1
run prehi0 MOGP_PRE_HI=0
run prehi1_la64 MOGP_LOOKAHEAD=64
