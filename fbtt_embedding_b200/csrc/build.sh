#!/usr/bin/env bash
# Builds libttb.so (the C-ABI CUDA library, include/ttb.h) for sm_100a, in-tree.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$ROOT/fbtt_embedding_b200/lib"
mkdir -p "$OUT" "$OUT/obj"
NVCC=${NVCC:-nvcc}
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I$ROOT/include -I$HERE ${TTB_NVCC_EXTRA:-}"
pids=()
for src in "$HERE"/*.cu; do
  obj="$OUT/obj/$(basename "${src%.cu}").o"
  stale=""
  for dep in "$src" "$HERE"/*.cuh "$ROOT/include/ttb.h" "$HERE/build.sh"; do
    if [ ! -f "$obj" ] || [ "$dep" -nt "$obj" ]; then stale=1; fi
  done
  if [ -n "$stale" ] || [ -n "${FORCE:-}" ]; then
    $NVCC $FLAGS -c "$src" -o "$obj" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/libttb.so" "$OUT"/obj/*.o -lcudart_static -lpthread -ldl -lrt
echo "[ttb] built $OUT/libttb.so"
