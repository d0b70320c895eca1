#!/bin/bash
# Builds libsaunet_b200.so in-tree for sm_100a (explicit nvcc; no JIT cache).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libsaunet_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr ${SAUNET_NVCC_EXTRA}"
mkdir -p "$HERE/_obj"
pids=()
for f in "$HERE"/*.cu; do
  o="$HERE/_obj/$(basename "${f%.cu}").o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ "$HERE/common.cuh" -nt "$o" ] || [ "$HERE/tc_common.cuh" -nt "$o" ] || [ "$HERE/../../include/saunet_b200.h" -nt "$o" ]; then
    $NVCC $FLAGS -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait "$p"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "$HERE"/_obj/*.o -lcuda
echo "built $OUT"
