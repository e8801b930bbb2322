#!/bin/bash
# Variant builds of chamfer.cu (compile-time switches) timed with scripts/time_knn.py; objects under /tmp (nothing ships).
set -e
cd "$(dirname "$0")/.."
PKG=mesh_r-cnn_computer_vision_project_b200
python -m meshrcnn_b200.build > /dev/null
mkdir -p /tmp/mrb_variants
for v in "$@"; do
  tag=$(echo "$v" | tr -d ' =.-' )
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I include $v -c $PKG/csrc/chamfer.cu -o /tmp/mrb_variants/chamfer_$tag.o 2>/dev/null
  objs=$(ls $PKG/build/*.o | grep -v "/chamfer.o")
  nvcc -shared -o /tmp/mrb_variants/k_$tag.so $objs /tmp/mrb_variants/chamfer_$tag.o -gencode arch=compute_100a,code=sm_100a -lcuda
  echo "== variant $v"
  MRB_LIB_PATH=/tmp/mrb_variants/k_$tag.so python scripts/time_knn.py 10 2 2>&1 | tail -1
done
