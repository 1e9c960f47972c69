# Builds the product library (hand-written sm_100a kernels + C ABI + C++ host API), the test oracle
# and, when /root/reference is present, the reference's own CUDA build (oracle/_ref).
NVCC      ?= /usr/local/cuda/bin/nvcc
HOSTCXX   := g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
# -fmad=false: every fused multiply-add in the kernels is an explicit __fmaf_rn (bit parity with the
# reference, see csrc/kernels_solve.cu); the compiler must not add its own.
NVFLAGS   := $(ARCH) -ccbin $(HOSTCXX) -std=c++17 -O3 -lineinfo -fmad=false -Xcompiler -fPIC,-O2,-fno-fast-math -Iinclude
PKG       := cuda_flow3d_b200
CSRC      := $(PKG)/csrc
LIB       := $(PKG)/libflow3d_b200.so
MGPU_LIB  := $(PKG)/libflow3d_b200_mgpu.so
CU_SRCS   := $(CSRC)/flow3d_cabi.cu $(CSRC)/kernels_solve.cu $(CSRC)/kernels_sweep_tma.cu $(CSRC)/kernels_pyramid.cu \
             $(CSRC)/kernels_warp.cu $(CSRC)/kernels_median.cu $(CSRC)/kernels_synth.cu \
             $(CSRC)/kernels_diag.cu
CU_OBJS   := $(CU_SRCS:.cu=.o)
HOST_SRCS := $(wildcard $(PKG)/host/*.cpp)
HOST_OBJS := $(HOST_SRCS:.cpp=.o)

all: $(LIB) $(MGPU_LIB) oracle apps

$(CSRC)/kernels_median.o: $(CSRC)/median_net_27.inc $(CSRC)/median5_sort25.inc

$(CSRC)/median5_sort25.inc: scripts/gen_median5_pair.py
	python3 scripts/gen_median5_pair.py $(CSRC)

$(CSRC)/median_net_%.inc: scripts/gen_median_network.py
	python3 scripts/gen_median_network.py $* $@

%.o: %.cu $(CSRC)/common.cuh $(CSRC)/solve_args.cuh include/flow3d_c.h
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; exit 1)

$(PKG)/host/%.o: $(PKG)/host/%.cpp
	$(HOSTCXX) -std=c++17 -O2 -fPIC -fno-fast-math -Iinclude -I$(PKG)/host -c $< -o $@

$(LIB): $(CU_OBJS) $(HOST_OBJS)
	$(NVCC) $(ARCH) -ccbin $(HOSTCXX) -shared -o $@ $^ -ldl

# multi-GPU: the z-sharded solver (host C++ over the C ABI above + NCCL); a separate library so that the
# single-GPU library carries no NCCL dependency
$(MGPU_LIB): $(CSRC)/sharded_solver.cu include/flow3d_mgpu_c.h include/flow3d_c.h $(LIB)
	$(NVCC) $(ARCH) -ccbin $(HOSTCXX) -std=c++17 -O2 -lineinfo -Xcompiler -fPIC,-O2,-fno-fast-math,-ffp-contract=off \
	  -Iinclude -shared -o $@ $< -L$(PKG) -lflow3d_b200 -lnccl -Xlinker -rpath -Xlinker '$$ORIGIN'

oracle:
	$(MAKE) -C oracle liboracle.so

# example programs against the reference-shaped C++ host API
APPS := build/flow3d_cli build/example_main build/flow3d_synth
apps: $(APPS)
# standalone input generator for bench.py's reference arm: the synth kernel compiled in, NOT linked to $(LIB)
build/flow3d_synth: apps/flow3d_synth.cu $(CSRC)/kernels_synth.cu $(CSRC)/common.cuh
	@mkdir -p build
	$(NVCC) $(ARCH) -ccbin $(HOSTCXX) -std=c++17 -O3 -fmad=false -Iinclude -I$(CSRC) $< -o $@
build/%: apps/%.cpp $(LIB)
	@mkdir -p build
	$(HOSTCXX) -std=c++17 -O2 -Iinclude $< -o $@ -L$(PKG) -lflow3d_b200 -Wl,-rpath,'$$ORIGIN/../$(PKG)'

ref:
	bash oracle/build_ref.sh

clean:
	rm -f $(CU_OBJS) $(HOST_OBJS) $(CSRC)/*.ptxas.log $(LIB) $(MGPU_LIB) $(APPS)
	$(MAKE) -C oracle clean

.PHONY: all oracle ref clean apps
