#!/bin/bash
# Builds libtmb variants that differ in tmb_fourier.cu's compile-time switches (for tools/r2_call31.sh; select one with
# TMB_LIB=build_variants/libtmb_<name>.so).  -DFW_NO_SKIP: no corner skip; -DFW_LINEAR_ROWS: rows of patches in grid order.
set -eu
cd "$(dirname "$0")/.."
make > /dev/null
mkdir -p build_variants
OBJS="tomobar_b200/csrc/tmb_capi.o tomobar_b200/csrc/tmb_elem.o tomobar_b200/csrc/tmb_proj.o tomobar_b200/csrc/tmb_tv.o tomobar_b200/csrc/tmb_tv_rof.o"
for v in "skip_cf:" "noskip_cf:-DFW_NO_SKIP" "skip_lin:-DFW_LINEAR_ROWS" "noskip_lin:-DFW_NO_SKIP -DFW_LINEAR_ROWS"; do
  name=${v%%:*}; flags=${v#*:}
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $flags -c tomobar_b200/csrc/tmb_fourier.cu -o build_variants/f_$name.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build_variants/libtmb_$name.so $OBJS build_variants/f_$name.o -lcufft
  echo "built build_variants/libtmb_$name.so"
done
