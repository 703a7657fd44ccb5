#!/bin/bash
# Build a variant of librdf_b200.so with extra nvcc flags (e.g. -DTB_U=8) into build/variants/<name>/librdf_b200.so;
# select it at run time with RDF_B200_LIB=<path>.  usage: tools/build_variant.sh <name> <nvcc flags...>
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
NAME="$1"; shift
OUT="$ROOT/build/variants/$NAME"
mkdir -p "$OUT/obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
pids=()
for f in "$ROOT"/3d-beats_b200/csrc/*.cu; do
  b="$(basename "$f" .cu)"
  $NVCC $FLAGS "$@" -c "$f" -o "$OUT/obj/$b.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared --cudart shared -Xlinker -rpath=/usr/local/cuda/lib64 -o "$OUT/librdf_b200.so" "$OUT"/obj/*.o -gencode arch=compute_100a,code=sm_100a
rm -rf "$OUT/obj"
echo "built $OUT/librdf_b200.so"
