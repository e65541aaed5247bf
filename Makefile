# Builds the in-tree shared library (C ABI in include/chemps2_b200.h) for sm_100a and the CPU checker in oracle/.
NVCC ?= nvcc
CXX ?= g++
CUDA_HOME ?= /usr/local/cuda
SRC := chemps2_b200/csrc
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
CXXFLAGS := -O2 -std=c++17 -fPIC -Wall -I$(CUDA_HOME)/include
OBJS := $(SRC)/b2_core.o $(SRC)/b2_pool.o $(SRC)/b2_ops.o $(SRC)/b2_sigma_plan.o $(SRC)/b2_heff.o $(SRC)/b2_compile.o $(SRC)/b2_update_plan.o $(SRC)/b2_sobject.o $(SRC)/b2_twodm.o $(SRC)/b2_capi.o $(SRC)/b2_capi_update.o $(SRC)/b2_capi_dmrg.o $(SRC)/b2_capi_twodm.o $(SRC)/b2_capi_davidson.o $(SRC)/b2_kernels.o $(SRC)/b2_blas1.o $(SRC)/b2_svd.o $(SRC)/b2_davidson.o
LIB := chemps2_b200/libchemps2_b200.so

CALLER := tests/cpp/_bin/dmrg_caller

all: $(LIB) oracle/libb2oracle.so $(CALLER)

# a C++ caller written against the mirror of the reference's public classes (include/chemps2_b200.hpp)
$(CALLER): tests/cpp/dmrg_caller.cpp include/chemps2_b200.hpp include/chemps2_b200.h $(LIB)
	mkdir -p tests/cpp/_bin
	$(CXX) -O2 -std=c++17 -Wall -Iinclude $< -o $@ -Lchemps2_b200 -lchemps2_b200 -Wl,-rpath,'$$ORIGIN/../../../chemps2_b200'

$(SRC)/b2_sigma_plan.o: $(SRC)/b2_sigma_plan_f4.inc $(SRC)/b2_sigma_plan_f5.inc
$(SRC)/b2_update_plan.o: $(SRC)/b2_update_plan_qx.inc
$(SRC)/%.o: $(SRC)/%.cpp $(wildcard $(SRC)/*.h) include/chemps2_b200.h
	$(CXX) $(CXXFLAGS) -c $< -o $@
$(SRC)/%.o: $(SRC)/%.cu $(wildcard $(SRC)/*.h)
	$(NVCC) $(NVFLAGS) -c $< -o $@
$(LIB): $(OBJS)
	$(NVCC) -shared -o $@ $(OBJS) -cudart static -lpthread -ldl -lrt
oracle/libb2oracle.so: oracle/plan_exec.c oracle/worklist_emul.cpp oracle/svd_block_emul.cpp $(wildcard $(SRC)/*.h) include/chemps2_b200.h
	gcc -O2 -fPIC -c oracle/plan_exec.c -o oracle/plan_exec.o
	$(CXX) -O2 -std=c++17 -fPIC -c oracle/worklist_emul.cpp -o oracle/worklist_emul.o
	$(CXX) -O2 -std=c++17 -fPIC -c oracle/svd_block_emul.cpp -o oracle/svd_block_emul.o
	$(CXX) -shared -o $@ oracle/plan_exec.o oracle/worklist_emul.o oracle/svd_block_emul.o

clean:
	rm -f $(SRC)/*.o oracle/*.o $(LIB) oracle/libb2oracle.so $(CALLER)
.PHONY: all clean
