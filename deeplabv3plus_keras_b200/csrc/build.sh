#!/usr/bin/env bash
# Build libdlv3p.so for sm_100a, in-tree (the .so travels to the GPU box with the repo snapshot).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libdlv3p.so"
BUILD="${HERE}/build"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v
       -I"${HERE}/../../include")
if [ "${DLV3P_DIAG:-0}" = "1" ]; then
  # diagnostics build (scripts/gemm_decompose.py, A/B switches): work-skipping switches compiled IN, separate library
  FLAGS+=(-DDLV3P_DIAG); OUT="${HERE}/../libdlv3p_diag.so"; BUILD="${HERE}/build_diag"
fi
mkdir -p "${BUILD}"
pids=()
for f in api dwconv dwconv_tma eltwise loss preprocess gemm_simt gemm_tcgen05; do
  "${NVCC}" "${FLAGS[@]}" -c "${HERE}/${f}.cu" -o "${BUILD}/${f}.o" > "${BUILD}/${f}.log" 2>&1 &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
if [ $rc -ne 0 ]; then cat "${BUILD}"/*.log | grep -v "^ptxas info" | head -80; exit 1; fi
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -shared -o "${OUT}" "${BUILD}"/*.o -lcudart
echo "built ${OUT}"
