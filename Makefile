# Builds the C-ABI library (sm_100a), the C++ host-mirror test and the CPU oracle without Python.
# `python __graft_entry__.py` does the same and is what the tests use.
NVCC      ?= nvcc
CXX       ?= g++
NVCCFLAGS ?= -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -cudart static
PKG       := bioshell_b200
LIB       := $(PKG)/libbioshell_align.so
CSRC      := $(wildcard $(PKG)/csrc/*.cu $(PKG)/csrc/*.cuh $(PKG)/csrc/*.inl) include/bioshell_align.h

.PHONY: all lib host oracle test clean
all: lib host oracle

lib: $(LIB)
$(LIB): $(CSRC)
	$(NVCC) $(NVCCFLAGS) -o $@ $(PKG)/csrc/bsa_api.cu

host: $(PKG)/host/test_host_mirror $(PKG)/host/test_clustering_host $(PKG)/host/test_bucket_host
$(PKG)/host/test_bucket_host: $(PKG)/host/test_bucket_host.cpp $(PKG)/host/bioshell_bucket.hpp $(PKG)/host/bioshell_seq.hpp $(LIB)
	$(CXX) -std=c++17 -O2 -Wall -o $@ $< -L$(PKG) -lbioshell_align -Wl,-rpath,'$$ORIGIN/..'
$(PKG)/host/test_clustering_host: $(PKG)/host/test_clustering_host.cpp $(PKG)/host/bioshell_clustering.hpp $(PKG)/host/bioshell_seq.hpp $(LIB)
	$(CXX) -std=c++17 -O2 -Wall -o $@ $< -L$(PKG) -lbioshell_align -Wl,-rpath,'$$ORIGIN/..'
$(PKG)/host/test_host_mirror: $(PKG)/host/test_host_mirror.cpp $(PKG)/host/bioshell_seq.hpp $(LIB)
	$(CXX) -std=c++17 -O2 -Wall -o $@ $< -L$(PKG) -lbioshell_align -Wl,-rpath,'$$ORIGIN/..'

oracle:
	$(MAKE) -C oracle

test:
	python -m pytest tests -q -m "not gpu"

clean:
	rm -f $(LIB) $(PKG)/host/test_host_mirror $(PKG)/host/test_clustering_host $(PKG)/host/test_bucket_host
	rm -rf oracle/_ref
