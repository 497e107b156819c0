#!/bin/bash
# A/B of kernel variants in one GPU call (development aid).
#   1. here (no GPU needed):  tools/variant_bench.sh build name1:"-DFLAG1" name2:"-DFLAG2 -DX=3" ...
#      builds a2d-shells_b200/lib/variants/liba2ds_<name>.so per variant (they travel with gpurun)
#   2. on the GPU box:        gpurun -- 'bash tools/variant_bench.sh run [nx]'
#      runs tools/quick_bench.py with the default library and with every variant (A2DS_LIB),
#      side by side in gpurun_out/variants.txt
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
VDIR="$ROOT/a2d-shells_b200/lib/variants"
if [ "$1" = build ]; then
  shift; mkdir -p "$VDIR"
  for v in "$@"; do
    name="${v%%:*}"; flags="${v#*:}"
    echo "== $name: $flags"
    /usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo \
      -Xcompiler -fPIC -shared $flags -Xptxas -v -o "$VDIR/liba2ds_$name.so" \
      "$ROOT"/a2d-shells_b200/csrc/a2ds.cu "$ROOT"/a2d-shells_b200/csrc/mesh_io.cpp \
      "$ROOT"/a2d-shells_b200/csrc/partition.cpp -lnccl 2>&1 |
      grep -E "Compiling entry|Used|spill" | grep -A2 k_assemble | grep -E "Compiling|Used|spill" |
      paste - - - | sed -E "s/.*function '([^']+)'.* ([0-9]+) bytes stack frame, ([0-9]+) bytes spill stores, ([0-9]+) bytes spill loads.*Used ([0-9]+) registers.*/  \5 regs, spill st \3 ld \4 B  \1/"
  done
elif [ "$1" = run ]; then
  nx="${2:-700}"; mkdir -p "$ROOT/gpurun_out"; out="$ROOT/gpurun_out/variants.txt"; : > "$out"
  cd "$ROOT"
  echo "== default" >> "$out"; timeout 120 python tools/quick_bench.py "$nx" >> "$out" 2>&1 || true
  for lib in "$VDIR"/liba2ds_*.so; do
    [ -e "$lib" ] || continue
    echo "== $(basename "$lib")" >> "$out"
    A2DS_LIB="$lib" timeout 120 python tools/quick_bench.py "$nx" >> "$out" 2>&1 || true
  done
  cat "$out"
else
  sed -n 2,9p "$0"
fi
