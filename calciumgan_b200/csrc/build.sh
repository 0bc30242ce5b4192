#!/usr/bin/env bash
# Build libcalciumgan_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${CG_OUT:-$HERE/../libcalciumgan_b200.so}"   # CG_OUT: side builds for same-box A/B runs (CG_LIB=...)
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -std=c++17 -O3 -lineinfo \
  -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function \
  ${CG_PTXAS_V:+-Xptxas -v} ${CG_TC_INSTRUMENT:+-DCG_TC_INSTRUMENT} \
   -shared -o "$OUT" "$HERE/cg_engine.cu"
echo "built $OUT"
