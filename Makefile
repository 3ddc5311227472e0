# libgsdr.so -- B200 (sm_100a) RX/TX readout DSP path.  In-tree build: the .so travels with the
# gpurun snapshot (git-ignored, not gpurun-ignored).
NVCC   ?= nvcc
ARCH   := -gencode arch=compute_100a,code=sm_100a
SRC    := gpu_sdr_b200/csrc
OUT    := gpu_sdr_b200/libgsdr.so
OBJDIR := build/obj
# hostlogic.cpp builds the taps: no fast-math, no FMA contraction (bit-exact with the reference's
# nvcc-default host flags).  Device code keeps nvcc's default -fmad=true like the reference.
NVFLAGS := -std=c++17 -O3 $(ARCH) -lineinfo -Xcompiler -fPIC,-O2,-ffp-contract=off,-fno-fast-math,-Wall
CU_SRCS := pfb_kernels pfb_wsp_p1 pfb_wsp_p2 pfb_wsp_p3 pfb_wsp_p4 chirp_kernels direct_kernels direct_tc_kernels direct_i8_kernels tones_kernels welch_kernels rx tx host
OBJS    := $(addprefix $(OBJDIR)/,$(addsuffix .o,$(CU_SRCS))) $(OBJDIR)/hostlogic.o

FEEDER := tests/cpp/bin/realtime_feeder

all: $(OUT) $(FEEDER)

# C++ real-time feeder (tests/cpp/realtime_feeder.cpp): plain g++, links the C-ABI only
$(FEEDER): tests/cpp/realtime_feeder.cpp include/gsdr.h $(OUT)
	@mkdir -p tests/cpp/bin
	g++ -O2 -std=c++17 -Wall -I include $< -o $@ $(OUT) -pthread -Wl,-rpath,'$$ORIGIN/../../../gpu_sdr_b200'

$(OBJDIR)/%.o: $(SRC)/%.cu $(SRC)/common.hpp $(SRC)/devmath.cuh $(SRC)/direct_common.cuh $(SRC)/packed_f32x2.cuh $(SRC)/pfb_fused.cuh include/gsdr.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; false)

$(OBJDIR)/hostlogic.o: $(SRC)/hostlogic.cpp $(SRC)/common.hpp include/gsdr.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -x cu -c $< -o $@

$(OUT): $(OBJS)
	$(NVCC) -shared $(ARCH) -o $@ $^

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf build $(OUT)

.PHONY: all oracle clean
