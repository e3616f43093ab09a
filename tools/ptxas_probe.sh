#!/bin/bash
# tools/ptxas_probe.sh [-DFLAG ...]: registers / spills / stack of the headline kernel instantiation alone (seconds, no GPU)
cd "$(dirname "$0")/.."
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xptxas -v --expt-relaxed-constexpr "$@" \
  -cubin -o build/fused_only.cubin tools/fused_only.cu 2>&1 | grep -E "error|spill|Used|stack" | head
