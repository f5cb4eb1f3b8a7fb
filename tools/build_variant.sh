#!/bin/bash
# usage: tools/build_variant.sh NAME "-DSWTPG_X=1 ..."  -> build/variants/libswtpg_NAME.so (tuning aid; load with SWTPG_LIB=...)
set -e
NAME=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
  -shared -o build/variants/libswtpg_$NAME.so fdreadoutlibs_b200/csrc/swtpg_capi.cu fdreadoutlibs_b200/csrc/framegen_capi.cu build_stage_copy.o
