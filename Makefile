# Builds libtmb.so (the C-ABI product library, sm_100a only) in-tree.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xptxas -v
SRCS := $(wildcard tomobar_b200/csrc/*.cu)
OBJS := $(SRCS:.cu=.o)
LIB := tomobar_b200/libtmb.so

all: $(LIB)

%.o: %.cu tomobar_b200/csrc/tmb_common.h tomobar_b200/csrc/tmb_tv_fused.cuh tomobar_b200/csrc/tmb_tv_common.cuh include/tmb.h
	$(NVCC) $(NVFLAGS) -c $< -o $@

# tmb_tv.cu keeps the round-1 ROF code under other C names (see the note at its top)
tomobar_b200/csrc/tmb_tv.o: NVFLAGS += -Dtmb_rof_tv=tmb_rof_tv_r1 -Dtmb_rof_tv_iter=tmb_rof_tv_iter_r1

$(LIB): $(OBJS)
	$(NVCC) -shared $(ARCH) -o $@ $(OBJS) -lcufft

clean:
	rm -f $(OBJS) $(LIB)
