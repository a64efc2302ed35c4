#!/usr/bin/env bash
# Build the sm_100a engine library in-tree (travels to the GPU box with the snapshot).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=${PW_OUT:-p_winds_b200/libpwinds_b200.so}
HASH=$(python -c "import __graft_entry__ as g; print(g.source_hash())")
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=${PW_FMAD:-false} ${PW_DEFS:-} \
  -DPW_SRC_HASH="\"$HASH\"" \
  --expt-relaxed-constexpr --expt-extended-lambda -Xcompiler -fPIC -shared \
  ${PW_PTXAS_V:+-Xptxas -v} -o $OUT p_winds_b200/csrc/pw_engine.cu
echo "built $OUT"
