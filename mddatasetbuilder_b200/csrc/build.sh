#!/bin/bash
# Builds libmdb_b200.so (sm_100a only) in-tree next to the Python package.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libmdb_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2,-Wall,-Wno-unused-function,-pthread ${MDB_NVCC_EXTRA}"
mkdir -p "$HERE/build"
pids=()
for f in "$HERE"/*.cu "$HERE"/*.cpp; do
  b="$(basename "$f")"
  o="$HERE/build/${b%.*}.o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ "$HERE/common.cuh" -nt "$o" ] || [ "$HERE/geom.cuh" -nt "$o" ] || [ "$HERE/../../include/mdb_b200.h" -nt "$o" ]; then
    $NVCC $FLAGS -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait "$p"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "$HERE"/build/*.o -lpthread
echo "built $OUT"
