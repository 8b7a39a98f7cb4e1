#!/bin/bash
# usage: profiles/tools/build_variant.sh NAME [-DKNOB=VALUE ...]   -> scratch/variants/libmate_NAME.so (MATE-4v8-9 shape only)
NAME=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DMATE_DEV_SHAPE "$@" \
  -o /root/repo/scratch/variants/libmate_$NAME.so /root/repo/mate_b200/csrc/mate_b200.cu
