#!/bin/bash
# Diagnostic: builds the library with clock64() stamps in k_gemm_tc (-DMRB_TC_TIMELINE) and prints CTA 0's role timeline.
set -e
cd "$(dirname "$0")/.."
PKG=mesh_r-cnn_computer_vision_project_b200
python -m meshrcnn_b200.build > /dev/null
mkdir -p /tmp/mrb_variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I include -DMRB_TC_TIMELINE -c $PKG/csrc/gemm_tc.cu -o /tmp/mrb_variants/gemm_tc_tl.o 2>/dev/null
objs=$(ls $PKG/build/*.o | grep -v "/gemm_tc.o")
nvcc -shared -o /tmp/mrb_variants/g_tl.so $objs /tmp/mrb_variants/gemm_tc_tl.o -gencode arch=compute_100a,code=sm_100a -lcuda
MRB_LIB_PATH=/tmp/mrb_variants/g_tl.so python scripts/gemm_timeline.py "$@"
