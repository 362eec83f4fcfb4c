#!/bin/bash
# Build a variant of libd4b200.so with extra compiler defines for A/B timing (tools/ab.sh):
#   tools/build_variant.sh <name> [-DFLAG ...]   ->  build_ab/<name>.so
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
obj=/tmp/d4_variants/$name; mkdir -p $obj $root/build_ab
cd $root/tad_dftd4_b200/csrc
pids=()
for cu in *.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 "$@" -Xcompiler -fPIC -c $cu -o $obj/${cu%.cu}.o 2> $obj/${cu%.cu}.log &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -o $root/build_ab/$name.so $obj/*.o 2>/dev/null
cuobjdump --dump-resource-usage $obj/flavour_f64_e.o 2>/dev/null | grep -o "Li[0-9]*ELi[0-9]*ELi[0-9]*E\|REG:[0-9]*\|STACK:[0-9]*" | paste - - - | tr '\n' ';'; echo
