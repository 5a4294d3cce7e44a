#!/bin/bash
# usage: tools/mk.sh NAME [-D...]   -> axial_vs_b200/libaxvs_NAME.so
n=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DAXVS_BUILD_ID=\"$n\" "$@" -shared -o axial_vs_b200/libaxvs_$n.so axial_vs_b200/csrc/axvs.cu 2>&1 | grep -i "error\|ptxas" | head
