# Builds libraytracing_cuda.so (the product) for sm_100a, the CPU oracle and the CPU kernel-body harness.
NVCC ?= nvcc
PKG := opencl-raytracing_b200
CSRC := $(PKG)/csrc
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden
HDRS := $(wildcard $(CSRC)/*.h) $(CSRC)/kernels.cuh include/rtcuda.h

all: $(PKG)/libraytracing_cuda.so oracle/liboracle.so tests/hostsim/libhostsim.so

build/kernels.o: $(CSRC)/kernels.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

build/api.o: $(CSRC)/api.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(PKG)/libraytracing_cuda.so: build/kernels.o build/api.o
	$(NVCC) -shared -o $@ $^ -lcudart_static -lrt -lpthread -ldl

oracle/liboracle.so: oracle/oracle.cpp include/rtcuda.h
	$(MAKE) -C oracle liboracle.so

tests/hostsim/libhostsim.so: tests/hostsim/hostsim.cpp $(HDRS)
	g++ -O2 -std=c++17 -fPIC -shared -pthread -o $@ $<

clean:
	rm -rf build $(PKG)/libraytracing_cuda.so oracle/liboracle.so tests/hostsim/libhostsim.so
