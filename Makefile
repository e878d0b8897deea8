# Build of the B200-native drop-in for bssrdf/CUDA-Winograd's hot path.
#   make            -> cuda-winograd_b200/libwinograd_b200.so  (C-ABI, include/*.h) and ./Test (the reference's CLI)
#   make selftest   -> tools/selftest (developer check against an in-program FP64 convolution)
# The reference's own Makefile passes no -arch (Makefile:14-17 there); tcgen05 needs the arch-specific target below.
NVCC    ?= nvcc
ARCH    := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude
CSRC    := cuda-winograd_b200/csrc
LIB     := cuda-winograd_b200/libwinograd_b200.so
KSRCS   := $(CSRC)/winograd_kernels.cu $(CSRC)/wino_small_kernel.cu $(CSRC)/wino_tm_kernel.cu $(CSRC)/wino_ff_kernel.cu $(CSRC)/wino_ffw_kernel.cu $(CSRC)/one_kernels.cu $(CSRC)/wg_api.cu $(CSRC)/legacy_entry.cu
HDRS    := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh include/*.h)

all: $(LIB) Test

$(LIB): $(KSRCS) $(CSRC)/host_util.c $(HDRS)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(KSRCS) $(CSRC)/host_util.c

Test: $(CSRC)/Test.c $(LIB)
	$(NVCC) $(ARCH) -O2 -Iinclude -o $@ $(CSRC)/Test.c -Lcuda-winograd_b200 -lwinograd_b200 \
	    -Xlinker -rpath -Xlinker '$$ORIGIN/cuda-winograd_b200'

selftest: tools/selftest
tools/selftest: tools/selftest.cu $(LIB)
	$(NVCC) $(NVFLAGS) -o $@ tools/selftest.cu -Lcuda-winograd_b200 -lwinograd_b200 \
	    -Xlinker -rpath -Xlinker '$$ORIGIN/../cuda-winograd_b200'

clean:
	rm -f $(LIB) Test tools/selftest $(CSRC)/*.o

.PHONY: all clean selftest
