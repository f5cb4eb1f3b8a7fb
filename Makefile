# Builds everything in-tree (the .so files travel to the GPU box with the snapshot; they are git-ignored).
#   make lib      fdreadoutlibs_b200/libswtpg_b200.so   CUDA kernels + C ABI (include/swtpg.h), sm_100a only
#   make host     fdreadoutlibs_b200/libswtpg_host.so   C++ frame-processor shim above the C ABI (g++)
#   make oracle   oracle/liboracle.so and (where /root/reference exists) oracle/_ref/libswtpg_ref.so
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH = -gencode arch=compute_100a,code=sm_100a
NVFLAGS = $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v --expt-relaxed-constexpr
CSRC = fdreadoutlibs_b200/csrc
LIB = fdreadoutlibs_b200/libswtpg_b200.so

all: lib host oracle

lib: $(LIB)

$(LIB): $(CSRC)/swtpg_capi.cu $(CSRC)/framegen_capi.cu $(CSRC)/stage_copy.cpp $(CSRC)/swtpg_kernels.cuh $(CSRC)/swtpg_device.cuh $(CSRC)/framegen.h include/swtpg.h include/swtpg_framegen.h
	g++ -O2 -std=c++17 -Wall -fPIC -c -o build_stage_copy.o $(CSRC)/stage_copy.cpp
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/swtpg_capi.cu $(CSRC)/framegen_capi.cu build_stage_copy.o 2> build_ptxas.log || (cat build_ptxas.log; exit 1)
	@grep -E "error|warning: v" build_ptxas.log || true

# Host-side C++ mirror of the reference's frame processors (plain g++; links against the C ABI only)
HOST = fdreadoutlibs_b200/libswtpg_host.so
host: $(HOST)
$(HOST): fdreadoutlibs_b200/host/swtpg_host.cpp fdreadoutlibs_b200/host/swtpg_host.hpp include/swtpg.h $(LIB)
	g++ -O2 -std=c++17 -Wall -Wextra -fPIC -shared -pthread -o $@ fdreadoutlibs_b200/host/swtpg_host.cpp -Lfdreadoutlibs_b200 -lswtpg_b200 -Wl,-rpath,'$$ORIGIN'

oracle:
	$(MAKE) -C oracle all

clean:
	rm -f $(LIB) $(HOST) build_ptxas.log
	$(MAKE) -C oracle clean
.PHONY: all lib host oracle clean
