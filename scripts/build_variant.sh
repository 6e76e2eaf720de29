#!/bin/bash
# usage: scripts/build_variant.sh <name> [nvcc flags for evolve.cu ...]  ->  build/var/var_<name>.so
# A kernel-variant library for A/B measurements (FB200_LIB=build/var/var_<name>.so python bench.py ...): only evolve.cu is
# recompiled with the extra flags; the other objects are built once into build/obj (rebuilt when a source is newer).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; shift
mkdir -p $ROOT/build/obj $ROOT/build/var
cd $ROOT/freddi_b200/csrc
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off"
for f in capi.cu fields.cu evolve_big.cu fields_big.cu resolve.cpp star.cpp; do
  o=$ROOT/build/obj/${f%.*}.o
  if [ ! -f $o ] || [ -n "$(find . ../../include -newer $o \( -name '*.cu' -o -name '*.cuh' -o -name '*.h' -o -name '*.hpp' -o -name '*.cpp' \) | head -1)" ]; then
    $NV -c $f -o $o &
  fi
done
$NV "$@" -c evolve.cu -o $ROOT/build/obj/evolve_$name.o &
wait
$NV -shared -o $ROOT/build/var/var_$name.so $ROOT/build/obj/evolve_$name.o $ROOT/build/obj/{capi,fields,evolve_big,fields_big,resolve,star}.o
echo built build/var/var_$name.so
