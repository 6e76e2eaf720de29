#!/bin/bash
# usage: scripts/build_variant.sh <name> [nvcc flags for the kernel translation units ...]  ->  build/var/var_<name>.so
# A kernel-variant library for A/B measurements (FB200_LIB=build/var/var_<name>.so python bench.py ...): evolve*.cu / fields*.cu
# are recompiled with the extra flags, the host objects are shared with nothing (freddi_b200/build.py keeps them per variant).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; shift
cd $ROOT && python -m freddi_b200.build --out build/var/var_$name.so -- "$@"
