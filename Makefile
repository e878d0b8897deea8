# Build of the B200-native drop-in for bssrdf/CUDA-Winograd's hot path.
#   make            -> cuda-winograd_b200/libwinograd_b200.so  (C-ABI, include/*.h) and ./Test (the reference's CLI)
#   make dev        -> tools/libwinograd_b200_dev.so (developer build: A/B knobs, ablation and superseded kernels)
#   make examples   -> examples/shard_batch (plain-C multi-GPU example: one host thread per GPU over the C-ABI)
#   make selftest   -> tools/selftest (developer check against an in-program FP64 convolution)
# The reference's own Makefile passes no -arch (Makefile:14-17 there); tcgen05 needs the arch-specific target below.
NVCC    ?= nvcc
ARCH    := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude
CSRC    := cuda-winograd_b200/csrc
LIB     := cuda-winograd_b200/libwinograd_b200.so
KSRCS   := $(CSRC)/winograd_kernels.cu $(CSRC)/wino_small_kernel.cu $(CSRC)/wino_ff_kernel.cu $(CSRC)/wino_ffw_kernel.cu $(CSRC)/conv3x3_direct_kernel.cu $(CSRC)/one_kernels.cu $(CSRC)/conv1x1_t_kernel.cu $(CSRC)/probe_kernels.cu $(CSRC)/wg_api.cu $(CSRC)/legacy_entry.cu
# developer build: + the superseded half-fold kernel, the ablation / CTA-pair instantiations and the WG_* environment knobs
DEVLIB  := tools/libwinograd_b200_dev.so
DEVSRCS := $(KSRCS) $(CSRC)/wino_tm_kernel.cu
HDRS    := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh include/*.h)

all: $(LIB) Test

# one object per source so that `make -j` compiles the kernels in parallel
OBJDIR  := build
OBJS    := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(KSRCS)) $(OBJDIR)/host_util.o
DEVOBJS := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/dev_%.o,$(DEVSRCS)) $(OBJDIR)/host_util.o

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c -o $@ $<
$(OBJDIR)/dev_%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -DWG_DEV_BUILD -c -o $@ $<
$(OBJDIR)/host_util.o: $(CSRC)/host_util.c $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c -o $@ $<

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS)

dev: $(DEVLIB)
$(DEVLIB): $(DEVOBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(DEVOBJS)

Test: $(CSRC)/Test.c $(LIB)
	$(NVCC) $(ARCH) -O2 -Iinclude -o $@ $(CSRC)/Test.c -Lcuda-winograd_b200 -lwinograd_b200 \
	    -Xlinker -rpath -Xlinker '$$ORIGIN/cuda-winograd_b200'

examples: examples/shard_batch
examples/shard_batch: examples/shard_batch.c $(LIB)
	$(NVCC) $(ARCH) -O2 -Iinclude -o $@ examples/shard_batch.c -Lcuda-winograd_b200 -lwinograd_b200 -lpthread \
	    -Xlinker -rpath -Xlinker '$$ORIGIN/../cuda-winograd_b200'

selftest: tools/selftest
tools/selftest: tools/selftest.cu $(DEVLIB)
	$(NVCC) $(NVFLAGS) -o $@ tools/selftest.cu -Ltools -lwinograd_b200_dev \
	    -Xlinker -rpath -Xlinker '$$ORIGIN'

clean:
	rm -rf $(LIB) $(DEVLIB) Test tools/selftest examples/shard_batch $(OBJDIR)

.PHONY: all clean selftest dev examples
