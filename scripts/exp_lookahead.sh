# (historical: MOGP_LAUNCH_AFTER / MOGP_REPLAY were removed when the cached columns became maintained)
run() { name=$1; shift; env "$@" python bench.py --steps 5 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2_exp_$name.json 2> gpurun_out/r2_exp_$name.err; }
run la40 MOGP_LOOKAHEAD=40
run la56 MOGP_LOOKAHEAD=56
run la64 MOGP_LOOKAHEAD=64
run la64np MOGP_LOOKAHEAD=64 MOGP_PRUNE=0
run after1 MOGP_LAUNCH_AFTER=1
run zig MOGP_ZIGZAG=1
