#!/bin/bash
# Builds diagnostic variants of the library (see the MRB_DIAG_* switches in csrc/gemm_tc.cu) into
# <package>/build/variants/<name>.so; select one at run time with MRB_LIB_PATH=...
set -e
cd "$(dirname "$0")/.."
PKG=mesh_r-cnn_computer_vision_project_b200
OUT=$PKG/build/variants
mkdir -p $OUT
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I include"
build() {  # name, extra flags
  nvcc $FLAGS $2 -c $PKG/csrc/gemm_tc.cu -o $OUT/gemm_tc_$1.o
  objs=$(ls $PKG/build/*.o | grep -v gemm_tc.o)
  nvcc -shared -o $OUT/$1.so $objs $OUT/gemm_tc_$1.o -gencode arch=compute_100a,code=sm_100a -lcuda
}
build noload "-DMRB_DIAG_NOLOAD" &
build nomma "-DMRB_DIAG_NOMMA" &
build nostore "-DMRB_DIAG_NOSTORE" &
build pf3 "-DMRB_TC_PREFETCH=3" &
build pf4 "-DMRB_TC_PREFETCH=4" &
wait
ls -la $OUT/*.so
