run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2_exp10_$name.json 2> gpurun_out/r2_exp10_$name.err; }
run kc8
run kc16 MOGP_RESCORE_KC=16
run kc0 MOGP_RESCORE_KC=0
run kc4 MOGP_RESCORE_KC=4
