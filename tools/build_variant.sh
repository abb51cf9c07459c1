#!/usr/bin/env bash
# Builds a variant of libdiasss_b200.so with extra preprocessor definitions, for A/B runs on the GPU box:
#   tools/build_variant.sh w4 -DDSX_FAST_WARPS=4      ->  diasss_b200/variants/libdiasss_b200_w4.so
#   DSX_LIB=diasss_b200/variants/libdiasss_b200_w4.so python bench.py ...
set -euo pipefail
name="$1"; shift
root="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
src="$root/diasss_b200/csrc"
out="$root/diasss_b200/variants"
bld="$src/build/variant_$name"
mkdir -p "$out" "$bld"
flags="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-Wall,-Wno-unused-function $*"
pids=()
for f in plan pyramid fast quadtree describe match frameprep kpspairs peer io capi; do
    /usr/local/cuda/bin/nvcc $flags -c "$src/$f.cu" -o "$bld/$f.o" &
    pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$out/libdiasss_b200_$name.so" "$bld"/*.o
echo "built $out/libdiasss_b200_$name.so"
