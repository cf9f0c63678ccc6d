# Convenience targets; the driver itself calls __graft_entry__.build() / pytest / bench.py directly.
PY ?= python

.PHONY: build test test-gpu bench bench-reference fuzz properties sanitize sanitize-host sanitize-oracle clean

build:            ## nvcc (sm_100a) + g++: libpdlp_b200.so, bin/pdlp_solve, the CPU checker
	$(PY) -c "import __graft_entry__ as g; g.build()"

test: build       ## CPU suite: oracle vs the reference's known answers, host logic, formats, C-ABI symbols
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu: build   ## the CUDA path through the C ABI against the oracle (needs a B200)
	$(PY) -m pytest tests -q -m gpu

bench: build      ## one JSON line (C2 on one GPU)
	$(PY) bench.py

bench-reference:  ## the CPU arm of the same metric
	$(PY) bench.py --impl reference

fuzz:             ## ASan + UBSan mutation fuzzer of the format layer (ITER=200000 make fuzz)
	tools/fuzz_formats.sh $${ITER:-200000}

properties:       ## long hypothesis campaign (EXAMPLES=3000 make properties)
	PDLP_B200_PROPERTY_EXAMPLES=$${EXAMPLES:-3000} $(PY) -m pytest tests/test_proto_codec_properties.py tests/test_host_logic_properties.py -q

sanitize:         ## compute-sanitizer passes on the small problems (needs a GPU)
	tools/sanitize.sh

sanitize-host:    ## the library's host side under ASan + UBSan through its host-only entry points (no GPU)
	tools/asan_host.sh

sanitize-oracle:  ## the multithreaded CPU checker under ThreadSanitizer, then ASan + UBSan (no GPU)
	tools/tsan_oracle.sh

clean:
	rm -rf build or-tools_b200/lib or-tools_b200/bin oracle/libpdlp_oracle.so
