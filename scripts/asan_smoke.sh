#!/bin/bash
# Host-side AddressSanitizer run of the library's threaded and single-threaded refit paths on a GPU box:
# builds libmogp_b200 with -fsanitize=address into build/ and runs __graft_entry__.smoke() with both optimisers.
# (CUDA needs protect_shadow_gap=0; leak detection off: the CUDA runtime keeps its allocations until exit.)
set -u
cd "$(dirname "$0")/.."
mkdir -p build
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O1 -g -std=c++17 \
  -Xcompiler -fPIC,-fsanitize=address,-fno-omit-frame-pointer -shared -o build/libmogp_asan.so mogp_b200/csrc/engine.cu || exit 1
ASAN=$(gcc -print-file-name=libasan.so)
for opt in native scipy; do
  echo "== smoke() under host ASAN, MOGP_OPTIMIZER=$opt"
  LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0:abort_on_error=0 MOGP_B200_LIB=$PWD/build/libmogp_asan.so \
    MOGP_OPTIMIZER=$opt timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -25
  echo "rc=$?"
done
