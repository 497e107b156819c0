# Convenience targets; the authoritative build is __graft_entry__.build() (same commands).
PY ?= python
LIBDIR := a2d-shells_b200/lib

lib:            ## CUDA library for sm_100a (nvcc) -> $(LIBDIR)/liba2ds_b200.so
	$(PY) a2d-shells_b200/build.py

oracle:         ## test infrastructure: plain-C restatement (+ the unmodified reference if present)
	$(MAKE) -C oracle all

example: lib    ## C++ driver without TACS: deck -> device Kmat / Gmat / residual
	g++ -std=c++11 -O2 -Iinclude examples/cylinder_buckling_assembly.cpp \
	    -L$(LIBDIR) -la2ds_b200 -Wl,-rpath,$(CURDIR)/$(LIBDIR) -o examples/cylinder_buckling_assembly

test-cpu:       ## everything that runs without a GPU
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu:       ## parity suite through the C ABI (needs a B200)
	$(PY) -m pytest tests -x -q -m gpu

bench:
	$(PY) bench.py --steps 10 --warmup 3

.PHONY: lib oracle example test-cpu test-gpu bench
