#!/usr/bin/env bash
# Build libphendiff_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${here}/../libphendiff_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"${NVCC}" -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC -shared \
  ${PD_NVCC_EXTRA:-} \
  -o "${out}" \
  "${here}/pd_api.cu" "${here}/pd_kernels_simt.cu" "${here}/pd_conv_tc.cu" "${here}/pd_conv_halo.cu" "${here}/pd_attn_mma.cu" "${here}/pd_attn_tc.cu" "${here}/pd_attn_tc2.cu"
echo "built ${out}"
