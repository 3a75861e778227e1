#!/bin/bash
# A/B builds of the library with other compile-time geometry: bash tools/build_variant.sh <name> -DPSCL_DICT_NT=384 ...
# -> build/libpscl_<name>.so (use with PSCL_LIB_PATH, see tools/gpu_variants.sh)
NAME=$1; shift
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -Xcompiler -fno-strict-aliasing \
  -diag-suppress 550 -Xptxas -v "$@" -o build/libpscl_$NAME.so popscle_b200/csrc/popscle_b200.cu 2> build/$NAME.ptxas.log
echo "built build/libpscl_$NAME.so: $?"
grep -E "k_demux_defaultILi8" -A2 build/$NAME.ptxas.log | grep -E "Compiling|registers|spill" | sed 's/ptxas info    : //' | paste - - - | awk '{print $4, $0}' | cut -c1-400 | sed 's/Compiling entry function//' | awk '{print}' | cut -d"'" -f2,3 | head -8
