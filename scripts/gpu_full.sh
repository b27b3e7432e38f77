#!/bin/bash
# Full validation in one gpurun call: all GPU tests (with durations), smoke, bench.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider --durations=15 > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$? (${SECONDS}s)"; tail -25 gpurun_out/pytest_all.log
SECONDS=0
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? (${SECONDS}s)"; tail -1 gpurun_out/smoke.log
SECONDS=0
timeout 600 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$? (${SECONDS}s)"; tail -2 gpurun_out/bench_full.err
python -c "
import json; d=json.load(open('gpurun_out/bench_full.json')); print(d['value'], d['ms_per_step'], d['phase_ms'], d['objects_encoded_per_s'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'], d['gpu_launches'])"
