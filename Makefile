# Builds libgfb200.so (C-ABI device layer + host front end bindings) for sm_100a, in tree.
CUDA ?= /usr/local/cuda
NVCC := $(CUDA)/bin/nvcc
# The system compiler: its libstdc++ is the shared one every other module of the process uses.
# (An environment CXX pointing at a toolchain with a static libstdc++ breaks iostreams inside a dlopen'ed library.)
CXX := /usr/bin/g++
PKG := graph_framework_b200
SRC := $(PKG)/csrc
LIB := $(PKG)/libgfb200.so
CXXFLAGS := -std=c++20 -O2 -g -fPIC -Wall -Wno-unused-function -I$(CUDA)/include -Iinclude
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC

all: $(LIB)

$(SRC)/skeleton_text.inc: $(SRC)/skeleton.cuh
	( printf 'R"GFBSKEL(' ; cat $< ; printf ')GFBSKEL"\n' ) > $@

# special.cuh with its coefficient table spliced in: NVRTC sees one self-contained text.
$(SRC)/special_text.inc: $(SRC)/special.cuh $(SRC)/wim_table.inc
	( printf 'R"GFBSPEC(' ; sed -e '/#include "wim_table.inc"/{r $(SRC)/wim_table.inc' -e 'd}' $< ; printf ')GFBSPEC"\n' ) > $@

build/kernels.o: $(SRC)/kernels.cu
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

build/runtime.o: $(SRC)/runtime.cpp $(SRC)/skeleton_text.inc $(SRC)/special_text.inc include/gfb200.h
	@mkdir -p build
	$(CXX) $(CXXFLAGS) -c $< -o $@

build/c_binding.o: $(SRC)/c_binding.cpp $(wildcard $(SRC)/graph/*.hpp) $(SRC)/special.cuh $(SRC)/wim_table.inc include/gfb200.h include/graph_c_binding.h include/gfb_rays.h
	@mkdir -p build
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(LIB): build/kernels.o build/runtime.o build/c_binding.o
	$(CXX) -shared -o $@ $^ -L$(CUDA)/lib64 -lcudart_static -lnvrtc -ldl -lrt -lpthread

clean:
	rm -rf build $(LIB) $(SRC)/skeleton_text.inc $(SRC)/special_text.inc

# Reference-style C++ programs built against the header-only front end + libgfb200.so.
TESTBIN := build/known_answers build/xrays_bench_b200
tests: $(TESTBIN)
build/known_answers: tests/cpp/known_answers.cpp $(wildcard $(SRC)/graph/*.hpp) $(LIB)
	$(CXX) -std=c++20 -O1 -Iinclude -I$(CUDA)/include $< -o $@ -L$(PKG) -lgfb200 -Wl,-rpath,'$$ORIGIN/../$(PKG)'
build/xrays_bench_b200: examples/xrays_bench.cpp $(wildcard $(SRC)/graph/*.hpp) $(LIB)
	$(CXX) -std=c++20 -O1 -Iinclude -I$(CUDA)/include $< -o $@ -L$(PKG) -lgfb200 -Wl,-rpath,'$$ORIGIN/../$(PKG)' -lpthread
