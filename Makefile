# Build of the B200-native Krylov hot path.  `make` builds everything in-tree:
#   slepc_b200/lib/libb200krylov.so  CUDA kernels + C ABI (include/b2k.h), sm_100a only
#   slepc_b200/lib/libb2kslepc.so    C host side mirroring SLEPc's BV/DS/ST/EPS/SVD API (include/b2kslepc.h)
#   oracle/_build/liboraclecpu.so    CPU BV/Mat plugin used ONLY as test oracle / timed CPU baseline
NVCC      ?= /usr/local/cuda/bin/nvcc
CC        ?= gcc
PYTHON    ?= python
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude -Islepc_b200/csrc
CFLAGS    := -O2 -g -fPIC -std=gnu11 -Wall -Wno-unused-function -Iinclude -Islepc_b200/host
OPENBLAS  := $(shell $(PYTHON) -c "import scipy,os,glob;print(os.path.realpath(glob.glob(os.path.join(os.path.dirname(scipy.__file__),'..','scipy.libs','libscipy_openblas*.so'))[0]))")
OPENBLAS_DIR := $(dir $(OPENBLAS))

LIBDIR    := slepc_b200/lib
KSRC      := $(wildcard slepc_b200/csrc/*.cu)
KHDR      := $(wildcard slepc_b200/csrc/*.h) include/b2k.h
HSRC      := $(wildcard slepc_b200/host/*.c)
HHDR      := $(wildcard slepc_b200/host/*.h) include/b2kslepc.h include/b2k.h
OSRC      := $(wildcard oracle/*.c)

ifneq ($(strip $(OSRC)),)
ORACLE_LIB := oracle/_build/liboraclecpu.so
endif
EXSRC     := $(wildcard examples/*.c)
EXBIN     := $(patsubst examples/%.c,examples/bin/%,$(EXSRC))
all: $(LIBDIR)/libb200krylov.so $(LIBDIR)/libb2kslepc.so $(ORACLE_LIB) $(EXBIN) baseline/libbase.so

# library baseline (cuBLAS / cuSPARSE in the reference's schedule) timed by bench.py next to the product; not part of it
baseline/libbase.so: baseline/libbase.cu
	$(NVCC) $(ARCH) -O3 -std=c++17 -Xcompiler -fPIC -shared -o $@ $< -lcublas -lcusparse -Xlinker -rpath=/usr/local/cuda/lib64

$(LIBDIR)/libb200krylov.so: $(KSRC) $(KHDR)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(KSRC) -ldl

$(LIBDIR)/libb2kslepc.so: $(HSRC) $(HHDR) $(LIBDIR)/libb200krylov.so
	$(CC) $(CFLAGS) -shared -o $@ $(HSRC) -L$(LIBDIR) -lb200krylov -Wl,-rpath,'$$ORIGIN' \
	    $(OPENBLAS) -Wl,-rpath,$(OPENBLAS_DIR) -lm -ldl

oracle/_build/liboraclecpu.so: $(OSRC) $(HHDR) $(LIBDIR)/libb2kslepc.so
	@mkdir -p oracle/_build
	$(CC) $(CFLAGS) -O3 -march=x86-64-v3 -fopenmp -B/usr/lib/gcc/x86_64-linux-gnu/13/ -shared -o $@ $(OSRC) -L$(LIBDIR) -lb2kslepc \
	    -Wl,-rpath,'$$ORIGIN/../../$(LIBDIR)' $(OPENBLAS) -Wl,-rpath,$(OPENBLAS_DIR) -lm

# the reference's tutorial programs against include/b2kslepc.h (run on the GPU only: B2KInitialize fails without one)
examples/bin/%: examples/%.c examples/exutil.h $(HHDR) $(LIBDIR)/libb2kslepc.so
	@mkdir -p examples/bin
	$(CC) $(CFLAGS) -Iexamples -o $@ $< -L$(LIBDIR) -lb2kslepc -lb200krylov -Wl,-rpath,'$$ORIGIN/../../$(LIBDIR)' -lm

examples: $(EXBIN)

check: all
	$(PYTHON) -m pytest tests -x -q -m "not gpu"

kernels: $(LIBDIR)/libb200krylov.so

clean:
	rm -f $(LIBDIR)/*.so oracle/_build/*.so examples/bin/* baseline/libbase.so

.PHONY: all clean kernels examples check
