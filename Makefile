# Build of the B200 engine.  `make` = the product (nvcc, sm_100a only).  `make simt` = the CPU unit-test build of the same kernel
# sources (tests only, see tests/simt/simt.h).  `make oracle` = the CPU oracle and, when /root/reference exists, oracle/_ref.
NVCC     ?= /usr/local/cuda/bin/nvcc
CXX      ?= g++
ARCH     := -gencode arch=compute_100a,code=sm_100a
EXTRA_NVFLAGS ?=
NVFLAGS  := $(ARCH) $(EXTRA_NVFLAGS) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden,-Wall,-Wno-unknown-pragmas --expt-relaxed-constexpr
CSRC     := mumemto_b200/csrc
CU       := $(wildcard $(CSRC)/*.cu)
HOSTCPP  := $(wildcard $(CSRC)/library_api.cpp)
OBJ      := $(patsubst $(CSRC)/%.cu,build/%.o,$(CU)) $(patsubst $(CSRC)/%.cpp,build/%.o,$(HOSTCPP))
LIB      := mumemto_b200/lib/libmumemto_b200.so
BIN      := mumemto_b200/bin/mumemto_exec mumemto_b200/bin/anchor_merge mumemto_b200/bin/extract_mums
HDRS     := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh include/*.h)

PYTHON   ?= python
PYEXT    := mumemto_b200/_mumemto_core$(shell $(PYTHON) -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
PYINC    := $(shell $(PYTHON) -c "import sysconfig, pybind11; print('-I' + sysconfig.get_paths()['include'], '-I' + pybind11.get_include())")

.PHONY: all lib simt oracle clean pymodule
all: lib $(BIN) pymodule
pymodule: $(PYEXT)
lib: $(LIB)

build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@
build/%.o: $(CSRC)/%.cpp $(HDRS)
	@mkdir -p build
	$(CXX) -O2 -std=c++17 -fPIC -fvisibility=hidden -Wall -Iinclude -c $< -o $@

$(LIB): $(OBJ)
	@mkdir -p mumemto_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart

CUDA_HOME ?= /usr/local/cuda
mumemto_b200/bin/mumemto_exec: $(CSRC)/cli_main.cpp $(CSRC)/multi_gpu.cpp $(CSRC)/multi_gpu.h $(LIB) $(HDRS)
	@mkdir -p mumemto_b200/bin
	$(CXX) -O2 -std=c++17 -Wall -Iinclude -I$(CUDA_HOME)/include $(CSRC)/cli_main.cpp $(CSRC)/multi_gpu.cpp -o $@ -Lmumemto_b200/lib -lmumemto_b200 \
	    -L$(CUDA_HOME)/lib64 -lcudart -lnccl -lz -lpthread -Wl,-rpath,'$$ORIGIN/../lib' -Wl,-rpath,$(CUDA_HOME)/lib64
mumemto_b200/bin/anchor_merge: $(CSRC)/anchor_merge_main.cpp $(LIB) $(HDRS)
	@mkdir -p mumemto_b200/bin
	$(CXX) -O2 -std=c++17 -Wall -Iinclude $< -o $@ -Lmumemto_b200/lib -lmumemto_b200 -Wl,-rpath,'$$ORIGIN/../lib'

mumemto_b200/bin/extract_mums: $(CSRC)/extract_mums_main.cpp
	@mkdir -p mumemto_b200/bin
	$(CXX) -O2 -std=c++17 -Wall $< -o $@ -lz

# pybind11 module with the reference's Python surface (python_bindings/src/mumemto_pybind.cpp) over the library ABI of include/mumemto.h
$(PYEXT): $(CSRC)/pybind_module.cpp include/mumemto.h $(LIB)
	$(CXX) -O2 -std=c++17 -fPIC -shared -fvisibility=hidden -Wall -Iinclude $(PYINC) $< -o $@ -Lmumemto_b200/lib -lmumemto_b200 -Wl,-rpath,'$$ORIGIN/lib'

# ---- tests only -------------------------------------------------------------------------------------------------------------
SIMT     := tests/simt/_build/libmumemto_b200_simt.so
simt: $(SIMT)
$(SIMT): $(CU) $(HOSTCPP) $(HDRS) tests/simt/simt.h tests/simt/simt.cpp
	@mkdir -p tests/simt/_build
	$(CXX) -O1 -g -std=c++17 -fPIC -shared -DMMB_SIMT -Itests/simt -Iinclude -Wall -Wno-unknown-pragmas -Wno-unused-function \
	    $(foreach f,$(CU),-x c++ $(f)) -x c++ $(HOSTCPP) tests/simt/simt.cpp -o $@

oracle:
	$(MAKE) -C oracle port ref

clean:
	rm -rf build mumemto_b200/lib mumemto_b200/bin mumemto_b200/_mumemto_core*.so tests/simt/_build
