# Builds the C-ABI shared library (sm_100a only) and the native self-check harness.
PKG   := vae-gan-based-model-for-image-generation-and-denoising_b200
CSRC  := $(PKG)/csrc
NVCC  ?= nvcc
ARCH  := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function
SRCS  := $(wildcard $(CSRC)/*.cu)
OBJS  := $(patsubst $(CSRC)/%.cu,build/obj/%.o,$(SRCS))
LIB   := $(PKG)/libvaegan_b200.so

all: $(LIB)

build/obj/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) include/vaegan_b200.h
	@mkdir -p build/obj
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS)

harness: $(LIB) tests/native/igemm_harness.cu
	@mkdir -p build
	$(NVCC) $(ARCH) -O2 -std=c++17 -o build/igemm_harness tests/native/igemm_harness.cu -L$(PKG) -lvaegan_b200 -Xlinker -rpath -Xlinker '$$ORIGIN/../$(PKG)'

# experiment build: the weight-gradient kernel's producer / issuer threads time their own phases (build/trace/)
trace:
	@mkdir -p build/trace
	for f in $(SRCS); do $(NVCC) $(NVFLAGS) -DVG_WGRAD_TRACE=1 -c $$f -o build/trace/$$(basename $$f .cu).o || exit 1; done
	$(NVCC) $(ARCH) -shared -o build/trace/libvaegan_b200.so build/trace/*.o
	$(NVCC) $(ARCH) -O2 -std=c++17 -DVG_WGRAD_TRACE=1 -o build/trace/igemm_harness tests/native/igemm_harness.cu -Lbuild/trace -lvaegan_b200 -Xlinker -rpath -Xlinker '$$ORIGIN'

# stand-alone hardware probes (tcgen05 issue rates, cluster launch rules, halo-tile descriptors): build/<name>
PROBES := umma_rate umma_rate2 umma_ring cluster_probe umma_halo
probes:
	@mkdir -p build
	for n in $(PROBES); do $(NVCC) $(ARCH) -O2 -std=c++17 -o build/$$n tests/native/$$n.cu || exit 1; done

# CPU-only check of the halo-tile plan (includes conv_api.cu, links the library's other objects; no CUDA call is made)
halo_plan_test: $(LIB) tests/native/halo_plan_test.cu
	@mkdir -p build
	$(NVCC) $(ARCH) -O1 -std=c++17 -Iinclude -o build/halo_plan_test tests/native/halo_plan_test.cu $(filter-out build/obj/conv_api.o,$(OBJS))

clean:
	rm -rf build $(LIB)

.PHONY: all harness trace probes halo_plan_test clean
