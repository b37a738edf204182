# Convenience targets; the Python entry points (__graft_entry__.build, sliceslice_rs_b200.build) do the
# same work and are what the tests and the bench call.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr
CSRC := sliceslice_rs_b200/csrc
OBJDIR := sliceslice_rs_b200/_obj
SRCS := capi.cu host_engine.cu capi_ctx.cu capi_exchange.cu scan_long.cu scan_ldg_u1.cu scan_ldg_u4.cu scan_tma_16.cu scan_tma_32.cu gen.cu batch.cu hist.cu hayset.cu service.cu
OBJS := $(addprefix $(OBJDIR)/,$(SRCS:.cu=.o))
LIB := sliceslice_rs_b200/libsliceslice_b200.so

all: lib oracle

lib: $(LIB)

$(OBJDIR)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/sliceslice_b200.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(ARCH) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -cudart static -lpthread

oracle:
	$(MAKE) -C oracle all

test-cpu: all
	python -m pytest tests -x -q -m "not gpu"

test-gpu: all
	python -m pytest tests -x -q -m gpu

clean:
	rm -rf $(OBJDIR) $(LIB) tests/cpp/_build
	$(MAKE) -C oracle clean

.PHONY: all lib oracle test-cpu test-gpu clean
