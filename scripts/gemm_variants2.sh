#!/bin/bash
set -e
cd "$(dirname "$0")/.."
PKG=mesh_r-cnn_computer_vision_project_b200
python -m meshrcnn_b200.build > /dev/null
mkdir -p $PKG/build/variants
for v in "-DMRB_TC_SINGLE_ACC_CHUNKS=0" "-DMRB_TC_SINGLE_ACC_CHUNKS=16" "-DMRB_TC_SINGLE_ACC_CHUNKS=0 -DMRB_TC_PREFETCH=3" "-DMRB_TC_SINGLE_ACC_CHUNKS=0 -DMRB_TC_PREFETCH=1"; do
  tag=$(echo "$v" | tr -d ' =-' )
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I include $v -c $PKG/csrc/gemm_tc.cu -o $PKG/build/variants/gemm_tc_$tag.o
  objs=$(ls $PKG/build/*.o | grep -v "/gemm_tc.o")
  nvcc -shared -o $PKG/build/variants/g_$tag.so $objs $PKG/build/variants/gemm_tc_$tag.o -gencode arch=compute_100a,code=sm_100a -lcuda
  echo "== variant $v"
  MRB_LIB_PATH=$PWD/$PKG/build/variants/g_$tag.so python scripts/time_gemm2.py
done
