#!/bin/bash
# builds the product library and the FB_PROFILE variant used by scripts/phase_profile.py
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $ROOT/build
(cd $ROOT/freddi_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -DFB_PROFILE "$@" -shared -o $ROOT/build/prof_default.so capi.cu evolve.cu fields.cu evolve_big.cu fields_big.cu resolve.cpp star.cpp)
(cd $ROOT && python -c "from freddi_b200 import build; build.build(force=True)")
echo built
