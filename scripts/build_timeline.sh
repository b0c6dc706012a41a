#!/bin/bash
# Diagnostic library with the in-kernel phase stamps of the fused convolution (-DDBOA_TIMELINE on conv_wide.cu only), linked next to
# the product objects into dynaboa_b200/build/libdboa_timeline.so; the product library is untouched.  Use with DBOA_LIB_PATH.
set -e
cd "$(dirname "$0")/.."
python -m dynaboa_b200.build

nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DDBOA_TIMELINE \
     -c dynaboa_b200/csrc/conv_wide.cu -o /tmp/conv_wide_tl.o
objs=$(ls dynaboa_b200/build/*.o | grep -v conv_wide.o)
nvcc -shared -o dynaboa_b200/build/libdboa_timeline.so $objs /tmp/conv_wide_tl.o -gencode arch=compute_100a,code=sm_100a -lcudart
ls -la dynaboa_b200/build/libdboa_timeline.so
