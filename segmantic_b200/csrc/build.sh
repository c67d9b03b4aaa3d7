#!/bin/bash
# Builds segmantic_b200/libsegmantic_b200.so for sm_100a (B200) in-tree.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libsegmantic_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC
       -Xcompiler -fvisibility=hidden)
mkdir -p "$HERE/build"
OBJS=()
PIDS=()
for src in "$HERE"/*.cu; do
  obj="$HERE/build/$(basename "${src%.cu}").o"
  stale=0
  for dep in "$src" "$HERE"/*.cuh "$HERE/../../include/segmantic_b200.h"; do
    if [[ ! -f "$obj" || "$dep" -nt "$obj" ]]; then stale=1; fi
  done
  if [[ $stale == 1 ]]; then
    rm -f "$obj"   # a failed compile must not leave a stale object for the link
    "$NVCC" "${FLAGS[@]}" ${SGM_PTXAS_V:+-Xptxas -v} -c "$src" -o "$obj" &
    PIDS+=($!)
  fi
  OBJS+=("$obj")
done
for pid in "${PIDS[@]}"; do
  wait "$pid" || { echo "build.sh: a compile failed" >&2; exit 1; }
done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT" "${OBJS[@]}"
echo "built $OUT"
