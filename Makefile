.SILENT:
# Builds lis_b200/_lib/liblis_b200.so : host C (lis.h API) + sm_100a CUDA kernels, one shared
# library with a C ABI.  `make` here or __graft_entry__.build().
NVCC     ?= /usr/local/cuda/bin/nvcc
CC       := /usr/bin/gcc
CUDA_INC := /usr/local/cuda/include
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -Iinclude
CFLAGS   := -O2 -g -fPIC -std=gnu11 -Wall -Wextra -Wno-unused-parameter -ffp-contract=off -Iinclude -Ilis_b200/csrc/host -I$(CUDA_INC)
OUT      := lis_b200/_lib
OBJ      := lis_b200/_build
KSRC     := $(wildcard lis_b200/csrc/kernels/*.cu)
HSRC     := $(wildcard lis_b200/csrc/host/*.c)
KOBJ     := $(patsubst lis_b200/csrc/kernels/%.cu,$(OBJ)/k_%.o,$(KSRC))
HOBJ     := $(patsubst lis_b200/csrc/host/%.c,$(OBJ)/h_%.o,$(HSRC))

# the reference's own drivers, compiled UNCHANGED from where they lie in the reference tree
# against include/ + liblis_b200.so (drop-in check).  Only where the reference tree exists; the
# binaries travel to the GPU box with the snapshot.
REF      ?= /root/reference
DRIVERS  := spmvtest1 spmvtest2 spmvtest2b spmvtest3 spmvtest3b spmvtest4 spmvtest5 test1 test2 test2b test3 test3b test3c test4 test5 etest1 etest2 etest3 etest4 etest5 etest5b etest6 etest7 test6 test7 getest1 getest5 getest5b
DRVBIN   := $(patsubst %,$(OUT)/drivers/%,$(DRIVERS))

.PHONY: all clean drivers
ifneq ($(wildcard $(REF)/test/spmvtest3.c),)
all: $(OUT)/liblis_b200.so $(OUT)/liblis_b200_shim.so drivers
drivers: $(DRVBIN)
$(OUT)/drivers/%: $(REF)/test/%.c $(OUT)/liblis_b200.so $(wildcard include/*.h)
	@mkdir -p $(OUT)/drivers
	$(CC) -O2 -DHAVE_CONFIG_H -Iinclude -o $@ $< -L$(OUT) -llis_b200 -Wl,-rpath,'$$ORIGIN/..' -lm
else
all: $(OUT)/liblis_b200.so $(OUT)/liblis_b200_shim.so
drivers:
	@echo "reference tree $(REF) not present: using prebuilt drivers (if any)"
endif

# the shared test driver (tests/shim/lis_shim.c, public API only) against this library
$(OUT)/liblis_b200_shim.so: tests/shim/lis_shim.c $(OUT)/liblis_b200.so $(wildcard include/*.h)
	$(CC) -O2 -g -fPIC -shared -Iinclude -o $@ tests/shim/lis_shim.c -L$(OUT) -llis_b200 -Wl,-rpath,'$$ORIGIN' -lm

$(OBJ)/k_%.o: lis_b200/csrc/kernels/%.cu lis_b200/csrc/kernels/common.cuh include/lis_b200_kernels.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(OBJ)/h_%.o: lis_b200/csrc/host/%.c $(wildcard lis_b200/csrc/host/*.h) $(wildcard include/*.h)
	@mkdir -p $(OBJ)
	$(CC) $(CFLAGS) -c $< -o $@

$(OUT)/liblis_b200.so: $(KOBJ) $(HOBJ)
	@mkdir -p $(OUT)
	$(NVCC) $(ARCH) -shared -Xlinker -Bsymbolic -o $@ $^ -cudart shared -lm -ldl -lpthread -lrt

clean:
	rm -rf $(OBJ) $(OUT)
