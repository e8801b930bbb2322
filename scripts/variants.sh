#!/bin/bash
# Generic variant runner: rebuilds ONE source file of the library with each set of -D flags, links it against the other
# objects of the in-tree build and runs a timing command with MRB_LIB_PATH pointing at the variant library.  Objects and
# libraries go to /tmp (nothing ships).
#   usage: variants.sh <file.cu under csrc/> "<timing command>" "<flags 1>" ["<flags 2>" ...]
#   e.g.   variants.sh gemm_tc.cu "python scripts/time_gemm2.py" "-DMRB_TC_PREFETCH=1" "-DMRB_TC_PREFETCH=3"
set -e
cd "$(dirname "$0")/.."
PKG=mesh_r-cnn_computer_vision_project_b200
SRC=$1; CMD=$2; shift 2
BASE=$(basename $SRC .cu)
python -m meshrcnn_b200.build > /dev/null
mkdir -p /tmp/mrb_variants
i=0
for v in "$@"; do
  i=$((i + 1))
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I include $v \
      -c $PKG/csrc/$SRC -o /tmp/mrb_variants/${BASE}_$i.o 2>/dev/null
  objs=$(ls $PKG/build/*.o | grep -v "/$BASE.o")
  nvcc -shared -o /tmp/mrb_variants/${BASE}_$i.so $objs /tmp/mrb_variants/${BASE}_$i.o -gencode arch=compute_100a,code=sm_100a -lcuda
  echo "== variant $v"
  MRB_LIB_PATH=/tmp/mrb_variants/${BASE}_$i.so $CMD 2>&1 | grep -v "^$" | tail -12
done
