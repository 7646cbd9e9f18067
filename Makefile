# phylign-b200 -- convenience targets (everything is plain python underneath)
PY ?= python

build:            ## CUDA library (nvcc, sm_100a) + CPU oracle (gcc)
	$(PY) -c "import __graft_entry__ as g; g.build()"

test:             ## CPU suite: oracle vs golden vectors, host logic, C-ABI surface
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu:         ## parity through the C ABI (needs a B200)
	$(PY) -m pytest tests -x -q -m gpu

smoke:            ## one small match on cuda:0 checked against the oracle
	$(PY) __graft_entry__.py smoke

bench:            ## BASELINE configs[2] on one GPU
	$(PY) bench.py

golden:           ## regenerate tests/golden (needs /root/reference)
	$(PY) tests/golden/make_golden.py

.PHONY: build test test-gpu smoke bench golden
