# Convenience targets; the driver uses __graft_entry__.build() which does the same.
NVCC ?= nvcc
NVFLAGS = -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -pthread -shared
LIB = gr_adsb_b200/csrc/libadsb_b200.so
SRC = gr_adsb_b200/csrc/adsb_api.cu
HDR = gr_adsb_b200/csrc/adsb_scan.cuh gr_adsb_b200/csrc/adsb_resolve.cuh gr_adsb_b200/csrc/adsb_device.cuh gr_adsb_b200/csrc/adsb_small.cuh include/adsb_b200.h

all: $(LIB) oracle/liboracle_adsb.so

$(LIB): $(SRC) $(HDR)
	$(NVCC) $(NVFLAGS) -o $@ $(SRC)

oracle/liboracle_adsb.so: oracle/adsb_oracle.c
	$(MAKE) -C oracle liboracle_adsb.so

test-cpu:
	python -m pytest tests -x -q -m "not gpu"

test-gpu:
	python -m pytest tests -x -q -m gpu

# the gpu parity tests against the host emulation of the CUDA sources (tests/emu): a logic check without a GPU
test-emu:
	ADSB_EMU=1 python -m pytest tests -x -q -m gpu

clean:
	rm -f $(LIB) oracle/liboracle_adsb.so tests/emu/libadsb_emu*.so

.PHONY: all test-cpu test-gpu test-emu clean
