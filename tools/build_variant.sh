#!/bin/bash
# tools/build_variant.sh NAME [-DFLAG ...]: cross-compiles build/libuwvkf_NAME.so from the product sources with extra defines
# (kernel tuning experiments; run with tools/phase_timing.py --lib build/libuwvkf_NAME.so)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
  -shared -o build/libuwvkf_$name.so slam-uwv_kalman_filters_b200/csrc/uwvkf.cu slam-uwv_kalman_filters_b200/csrc/small_filters.cu -lcudart 2>&1 | grep -E "error|spill" | head -5
echo built build/libuwvkf_$name.so
