# Builds libknockoffs_b200.so (sm_100a only) in-tree, plus the CPU oracle used by the tests.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v
PKG := knockoffs.jl_b200
SRC := $(PKG)/csrc/api.cu $(PKG)/csrc/solve.cu $(PKG)/csrc/ccd_stream.cu $(PKG)/csrc/ccd_block.cu $(PKG)/csrc/linalg.cu $(PKG)/csrc/pipeline.cu $(PKG)/csrc/group.cu $(PKG)/csrc/stats.cu $(PKG)/csrc/approx.cu $(PKG)/csrc/rep.cu $(PKG)/csrc/multi.cu $(PKG)/csrc/positive.cu $(PKG)/csrc/dgemm_tma.cu
HDR := $(wildcard $(PKG)/csrc/*.cuh) include/knockoffs_b200.h
OBJ := $(SRC:.cu=.o)
LIB := $(PKG)/libknockoffs_b200.so

all: $(LIB) oracle

%.o: %.cu $(HDR)
	$(NVCC) $(NVFLAGS) -c -o $@ $< 2> $@.ptxas.log || (cat $@.ptxas.log; exit 1)

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -cudart static

oracle:
	$(MAKE) -C oracle -s

clean:
	rm -f $(OBJ) $(PKG)/csrc/*.ptxas.log $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
