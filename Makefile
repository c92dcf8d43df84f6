# Top-level build: the product library (CUDA, sm_100a), the INT32 microbenchmark and the checkers.
#   make            -> crystals-kyber_b200/libmlkem_b200.so + build/microbench + oracle/
#   make lib        -> only the product library
NVCC    ?= /usr/local/cuda/bin/nvcc
ARCH    := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -cudart static
PKG     := crystals-kyber_b200
CSRC    := $(PKG)/csrc
LIB     := $(PKG)/libmlkem_b200.so

all: lib tools examples oracle

lib: $(LIB)

$(LIB): $(CSRC)/mlkem_b200.cu $(CSRC)/mlkem_kernels.cuh $(CSRC)/mlkem_device.cuh $(CSRC)/ml_kem_compat.inl $(CSRC)/mlkem_profile.inl $(CSRC)/sha3_compat.inl include/mlkem_b200.h include/ml_kem.h include/sha3.h
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/mlkem_b200.cu

# experiment build of the library (timing experiments that break the results on purpose: tools/time_encaps.py)
exp: build/libmlkem_b200_exp.so
build/libmlkem_b200_exp.so: $(CSRC)/mlkem_b200.cu $(CSRC)/mlkem_kernels.cuh $(CSRC)/mlkem_device.cuh
	mkdir -p build
	$(NVCC) $(NVFLAGS) -DMLKEM_B200_EXPERIMENT -shared -o $@ $(CSRC)/mlkem_b200.cu

tools: build/microbench build/keccak_bench build/coissue_bench

build/microbench: $(CSRC)/microbench.cu
	mkdir -p build
	$(NVCC) $(ARCH) -O3 -lineinfo -o $@ $<

build/coissue_bench: $(CSRC)/coissue_bench.cu
	mkdir -p build
	$(NVCC) $(ARCH) -O3 -lineinfo -o $@ $<

build/keccak_bench: $(CSRC)/keccak_bench.cu
	mkdir -p build
	$(NVCC) $(ARCH) -O3 -lineinfo -o $@ $<

# plain-C users of the batched ABI (gcc, no CUDA headers): compiled by `make` so that include/mlkem_b200.h stays valid C
examples: build/keyed_server
build/keyed_server: examples/keyed_server.c include/mlkem_b200.h $(LIB)
	mkdir -p build
	gcc -std=c99 -Wall -Wextra -O1 -Iinclude $< -L$(PKG) -lmlkem_b200 -Wl,-rpath,'$$ORIGIN/../$(PKG)' -o $@

oracle: lib
	$(MAKE) -C oracle
	$(MAKE) -C oracle drivers

clean:
	rm -f $(LIB) build/microbench build/keccak_bench build/coissue_bench
	$(MAKE) -C oracle clean

.PHONY: all lib exp tools examples oracle clean
