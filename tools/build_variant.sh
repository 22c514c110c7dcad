#!/bin/bash
# tools/build_variant.sh NAME [-DFLAG ...]: builds build/variants/libsrsue_gpu_NAME.so with extra compiler flags (tuning experiments)
set -e
name=$1; shift
cd "$(dirname "$0")/../srsue_b200/csrc"
make -s OBJDIR=../../build/obj_$name OUT=../../build/variants/libsrsue_gpu_$name.so NVFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --extended-lambda -Xcompiler -fPIC,-Wall -I../../include -I. $*" 2>&1 | grep -E "error|spill" || true
