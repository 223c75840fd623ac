#!/usr/bin/env bash
# ORACLE / TEST INFRASTRUCTURE ONLY.
# Compiles the reference's own raw CUDA kernels (tomobar/cuda_kernels/*.cu -- plain CUDA C++ that
# the reference JIT-compiles with NVRTC `-std=c++11`, cuda_kernels/__init__.py:12-30) from where
# they lie under /root/reference into oracle/_ref/*.cubin for sm_100a.  The cubins are loaded by
# tests/ref_kernels.py on the GPU box as the REAL reference for the TV / FBP-filter / USFFT
# kernels (kernel-level parity of libtmb against the reference's own code).  No reference source
# is copied into the repository; oracle/_ref/ is git-ignored (it still travels with gpurun).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC=/root/reference/tomobar/cuda_kernels
OUT="$HERE/_ref"
mkdir -p "$OUT"
for k in primal_dual_for_total_variation rudin_osher_fatemi_total_variation generate_filtersync fft_us_kernels; do
  if [ ! -f "$OUT/$k.cubin" ] || [ "$SRC/$k.cu" -nt "$OUT/$k.cubin" ]; then
    nvcc -cubin -std=c++11 -gencode arch=compute_100a,code=sm_100a -w -o "$OUT/$k.cubin" "$SRC/$k.cu"
  fi
done
ls -la "$OUT"
