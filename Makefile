# Convenience targets (everything is also reachable from Python: __graft_entry__.build / smoke, pytest, bench.py)
PY ?= python

.PHONY: build test test-gpu smoke bench bench-ref clean
build:            ## nvcc + g++: libb200fhe.so, libb200net.so, simulator, oracle, reference binaries (if the reference tree is mounted)
	$(PY) -c "import __graft_entry__ as g; g.build()"
test: build       ## CPU suite: oracle vs reference goldens, kernel simulator, host engine, front end, gloo sharding
	$(PY) -m pytest tests -q -m "not gpu"
test-gpu: build   ## on a B200: parity through the C ABI, netlists, command line, reference-side bindings
	$(PY) -m pytest tests -q -m gpu
smoke: build
	$(PY) -c "import __graft_entry__ as g; g.smoke()"
bench: build      ## BASELINE.json configs[1] on one GPU
	$(PY) bench.py --gpus 1
bench-ref: build  ## the unmodified TFHEpp on the host cores
	$(PY) bench.py --impl reference
clean:
	rm -f iyokan_b200/csrc/libb200fhe.so iyokan_b200/host/libb200net.so tests/sim/libbr_sim.so oracle/libtfhe_oracle.so
	rm -f scripts/microbench/pipes scripts/microbench/ntt_passes
	rm -rf oracle/_ref
