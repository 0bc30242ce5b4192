#!/usr/bin/env bash
# Build the CUDA library of another git revision into tools/libbase.so for same-box A/B runs:
#   tools/ab_build.sh HEAD~1 && gpurun -- 'bash tools/ab.sh CG_LIB=tools/libbase.so CG_X=1'
set -euo pipefail
REV="${1:-HEAD}"
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
TMP="$(mktemp -d)"
git -C "$ROOT" archive "$REV" calciumgan_b200/csrc include | tar -x -C "$TMP"
/usr/local/cuda/bin/nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC \
  -shared -o "$ROOT/tools/libbase.so" "$TMP/calciumgan_b200/csrc/cg_engine.cu"
rm -rf "$TMP"
echo "built tools/libbase.so from $REV"
