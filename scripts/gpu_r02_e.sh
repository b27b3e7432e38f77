#!/bin/bash
# round 2, GPU call E: full GPU suite, op-level bench (new thread-per-query ball query), C3 (DGCNN) at its named shape on one GPU
set -u
mkdir -p gpurun_out
SECONDS=0
python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^DEBUG\|^INFO" | tail -8; echo "pytest ${SECONDS}s"
SECONDS=0
python scripts/bench_ops.py > gpurun_out/r02_ops_bench.json 2> gpurun_out/r02_ops_bench.err; echo "ops rc=$? ${SECONDS}s"; tail -c 2500 gpurun_out/r02_ops_bench.json
SECONDS=0
python bench.py --config c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; echo "c3 rc=$? ${SECONDS}s"; tail -c 1500 gpurun_out/r02_bench_c3.json; tail -3 gpurun_out/r02_bench_c3.err
