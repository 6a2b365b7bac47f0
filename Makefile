# Builds ocrs_b200/libocrs_b200.so (CUDA kernels + C++ host library, C ABI in include/ocrs_b200.h)
# for sm_100a, and the oracle's C helpers.  Used by __graft_entry__.build().
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
CSRC      := ocrs_b200/csrc
BUILD     := build
CXXFLAGS  := -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function,-ffp-contract=off -lineinfo -Iinclude -I$(CSRC)
NVFLAGS   := $(ARCH) $(CXXFLAGS)
LIB       := ocrs_b200/libocrs_b200.so

# "exact" translation units: no FMA contraction on the device either
EXACT_CU  := $(CSRC)/image_kernels.cu $(CSRC)/ctc_beam.cu
FAST_CU   := $(filter-out $(EXACT_CU),$(wildcard $(CSRC)/*.cu))
CPP       := $(wildcard $(CSRC)/*.cpp)

OBJS := $(patsubst $(CSRC)/%.cu,$(BUILD)/%.o,$(EXACT_CU) $(FAST_CU)) $(patsubst $(CSRC)/%.cpp,$(BUILD)/%.cpp.o,$(CPP))
HDRS := $(wildcard $(CSRC)/*.h) $(wildcard include/*.h)

all: $(LIB)

$(BUILD):
	mkdir -p $(BUILD)

$(BUILD)/image_kernels.o: $(CSRC)/image_kernels.cu $(HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) -fmad=false -c $< -o $@

$(BUILD)/ctc_beam.o: $(CSRC)/ctc_beam.cu $(HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) -fmad=false -c $< -o $@

$(BUILD)/%.o: $(CSRC)/%.cu $(HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(BUILD)/%.cpp.o: $(CSRC)/%.cpp $(HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) -x cu -fmad=false -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lpthread

clean:
	rm -rf $(BUILD) $(LIB)

.PHONY: all clean
