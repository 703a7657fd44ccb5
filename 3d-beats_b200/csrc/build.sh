#!/bin/bash
# Builds librdf_b200.so (sm_100a only) in-tree, next to the Python package that loads it.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../rdf_b200/librdf_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
mkdir -p "$HERE/_obj"
pids=()
for f in rdf_capi rdf_eval rdf_layered rdf_meanshift rdf_synth rdf_train rdf_grouping rdf_tex rdf_frame rdf_points; do
  [ -f "$HERE/$f.cu" ] || continue
  if [ ! -f "$HERE/_obj/$f.o" ] || [ "$HERE/$f.cu" -nt "$HERE/_obj/$f.o" ] || [ -n "$(find "$HERE" "$HERE/../../include" -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$HERE/_obj/$f.o")" ]; then
    $NVCC $FLAGS ${RDF_NVCC_EXTRA} -c "$HERE/$f.cu" -o "$HERE/_obj/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
# --cudart shared: the static runtime would embed its whole symbol table in the .so; the shared one is already in every process
# that loads this library next to torch (rpath covers plain-C hosts)
$NVCC -shared --cudart shared -Xlinker -rpath=/usr/local/cuda/lib64 -o "$OUT" "$HERE"/_obj/*.o -gencode arch=compute_100a,code=sm_100a
echo "built $OUT"
