#!/bin/bash
# usage: tools/build_variant.sh NAME [-DFLAG ...]   -> build/libigtmpc_NAME.so (developer experiments; run with IGT_LIB=build/libigtmpc_NAME.so)
NAME=$1; shift
mkdir -p build
exec nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -shared \
    "$@" -o build/libigtmpc_$NAME.so igt_mpc_int_b200/csrc/igt_abi.cu
