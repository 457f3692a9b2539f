#!/bin/bash
# Build libmzb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="$HERE/../libmzb200.so"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
       -I"$ROOT/include" -I"$HERE" --fmad=true -Xptxas -v)
mkdir -p "$HERE/build"
objs=()
pids=()
for src in "$HERE"/*.cu; do
  obj="$HERE/build/$(basename "${src%.cu}").o"
  objs+=("$obj")
  if [[ ! -f "$obj" || "$src" -nt "$obj" || -n "$(find "$HERE" "$ROOT/include" \( -name '*.cuh' -o -name '*.h' \) -newer "$obj" -print -quit)" ]]; then
    "$NVCC" "${FLAGS[@]}" -c "$src" -o "$obj" > "$obj.log" 2>&1 &
    pids+=($!)
  fi
done
fail=0
for p in "${pids[@]:-}"; do [[ -z "$p" ]] || wait "$p" || fail=1; done
if [[ $fail -ne 0 ]]; then cat "$HERE"/build/*.log; exit 1; fi
"$NVCC" -shared -o "$OUT" "${objs[@]}" -lcudart
echo "built $OUT"
