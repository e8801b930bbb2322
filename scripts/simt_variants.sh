#!/bin/bash
# Variant builds of gemm_simt.cu (compile-time switches) timed with scripts/time_skinny.py; objects under /tmp.
set -e
cd "$(dirname "$0")/.."
PKG=mesh_r-cnn_computer_vision_project_b200
python -m meshrcnn_b200.build > /dev/null
mkdir -p /tmp/mrb_variants
for v in "$@"; do
  tag=$(echo "$v" | tr -d ' =.-' )
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I include $v -c $PKG/csrc/gemm_simt.cu -o /tmp/mrb_variants/simt_$tag.o 2>/dev/null
  objs=$(ls $PKG/build/*.o | grep -v "/gemm_simt.o")
  nvcc -shared -o /tmp/mrb_variants/s_$tag.so $objs /tmp/mrb_variants/simt_$tag.o -gencode arch=compute_100a,code=sm_100a -lcuda
  echo "== variant $v"
  MRB_LIB_PATH=/tmp/mrb_variants/s_$tag.so python scripts/time_skinny.py 2>&1 | grep skinny
done
