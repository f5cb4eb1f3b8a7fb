#!/bin/bash
# usage: tools/build_variant.sh NAME "-DSWTPG_X=1 ..."  -> build/variants/libswtpg_NAME.so (tuning aid; load with SWTPG_LIB=...)
# Only the kernel translation unit is rebuilt with the extra macros; the streaming engine and host utilities come from build/obj.
set -e
NAME=$1; shift
mkdir -p build/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
  -c -o build/variants/capi_$NAME.o fdreadoutlibs_b200/csrc/swtpg_capi.cu
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/libswtpg_$NAME.so build/variants/capi_$NAME.o \
  build/obj/swtpg_stream.o build/obj/swtpg_sort.o build/obj/swtpg_hostutil.o -lpthread
rm -f build/variants/capi_$NAME.o
