# convenience targets; the driver's entry points are __graft_entry__.py and bench.py
PY ?= python

.PHONY: build test test-gpu bench reference example clean

build:            ## nvcc for sm_100a (in-tree libchicdiff_b200.so) + the CPU oracle
	$(PY) __graft_entry__.py

test: build       ## CPU suite: oracle, golden table, ABI, glue, device math on the host, 2-rank gloo
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu: build   ## parity through the C ABI (needs a B200)
	$(PY) -m pytest tests -q -m gpu

bench: build      ## one JSON line (needs a B200)
	$(PY) bench.py

reference: build  ## the CPU restatement timed on this box's cores
	$(PY) bench.py --impl reference

example: build    ## the boundary from plain C
	gcc -std=c11 -Iinclude examples/c_abi_example.c -Lchicdiff_b200 -lchicdiff_b200 -Wl,-rpath,$(CURDIR)/chicdiff_b200 -lm -o c_abi_example

clean:
	rm -rf chicdiff_b200/csrc/_obj chicdiff_b200/libchicdiff_b200.so oracle/liboracle.so c_abi_example
