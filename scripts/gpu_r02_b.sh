#!/bin/bash
# round 2, GPU call B: new torch-op / QueryAndGroup tests first, then the whole GPU suite, smoke, and a short default bench
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_torch_ops.py -x -q 2>&1 | tail -15
SECONDS=0
python -m pytest tests -m gpu -x -q 2>&1 | tail -5; echo "pytest ${SECONDS}s"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
SECONDS=0
python bench.py --no-extra > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; echo "bench rc=$? ${SECONDS}s"; tail -c 1500 gpurun_out/r02_bench_b.json
