#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_all.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_fast4.json 2> gpurun_out/bench_fast4.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_fast4.json')); print(d['value'], d['ms_per_step'], d['phase_ms'], d['objects_encoded_per_s'], d['roofline']['frac'], d['e2e']['value'])"; tail -3 gpurun_out/bench_fast4.err
timeout 600 python scripts/bench_ops.py > gpurun_out/ops_bench.json 2> gpurun_out/ops_bench.err; echo "ops rc=$?"; tail -3 gpurun_out/ops_bench.err
