#!/bin/bash
# Occupancy variants of the split-input gather kernels (csrc/graphconv2.cu): -DGC2_FWD_MINB / -DGC2_BWD_MINB = CTAs per SM the
# kernels are compiled for (8 -> 32 registers, 6 -> 40, 5 -> 48).  Builds one library per variant next to the default one and
# runs scripts/bench_gather.py on each.    bash scripts/gather_variants.sh [grid]
set -e
cd "$(dirname "$0")/.."
PKG=mesh_r-cnn_computer_vision_project_b200
python -m meshrcnn_b200.build > /dev/null
mkdir -p $PKG/build/variants gpurun_out
for v in "8 6 6 1" "8 6 6 4" "8 6 6 16" "8 6 6 1000" "8 6 8 4" "8 5 6 4"; do
  set -- $v
  tag=f$1b$2p$3w$4
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I include \
       -DGC2_FWD_MINB=$1 -DGC2_BWD_MINB=$2 -DGC2_FWD_MINB_POS=$3 -DGC2_WAVES=$4 -c $PKG/csrc/graphconv2.cu -o $PKG/build/variants/graphconv2_$tag.o
  objs=$(ls $PKG/build/*.o | grep -v graphconv2.o)
  nvcc -shared -o $PKG/build/variants/$tag.so $objs $PKG/build/variants/graphconv2_$tag.o -gencode arch=compute_100a,code=sm_100a -lcuda
  echo "== variant fwd_minb=$1 bwd_minb=$2 fwd_minb_pos=$3 waves=$4"
  MRB_LIB_PATH=$PWD/$PKG/build/variants/$tag.so python scripts/bench_gather.py ${GRID:-48} 2>&1 | grep "_us"
done
