#!/bin/bash
# usage: tools/build_variant.sh <name> [-DMACRO=..]...   -> build/variants/libsde_<name>.so (for same-box A/B runs:
# SDE_LIB_PATH=build/variants/libsde_<name>.so python tools/ab_step.py)
set -e
cd "$(dirname "$0")/.."
name="$1"; shift
mkdir -p build/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -cudart static -diag-suppress 550 \
  "$@" -o build/variants/libsde_${name}.so simpledepthestimation_b200/csrc/*.cu
echo build/variants/libsde_${name}.so
