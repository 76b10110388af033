#!/bin/bash
# Perf bisects: builds csrc/build/variants/NAME.so = the shipped objects with the compat = physical walls group (lbm_step.cu, group 1)
# recompiled with extra -D flags.  Select it with LBM_B200_LIB=<path>.   usage: scripts/build_variant.sh NAME "-DFOO=1 ..."
set -e
cd "$(dirname "$0")/../pour_over_coffee_lbm_b200/csrc"
mkdir -p build/variants
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -O2 --expt-relaxed-constexpr"
$NV -fmad=false -DLBM_GROUP=1 -DLBM_STRICT_BUILD=1 $2 -c lbm_step.cu -o build/variants/$1_g1.o
OBJS=$(ls build/*.o | grep -v lbm_step_strict_g1.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/$1.so $OBJS build/variants/$1_g1.o -ldl
rm -f build/variants/$1_g1.o
echo built build/variants/$1.so
