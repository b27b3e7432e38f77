#!/bin/bash
# gen-2 matcher bring-up: intermediates diagnostic, fused tests, short bench (both generations)
set -u
mkdir -p gpurun_out
timeout 300 python scripts/debug_fused.py 256 > gpurun_out/debug_fused.log 2>&1; echo "debug rc=$?"; cat gpurun_out/debug_fused.log | tail -25
timeout 600 python -m pytest tests/test_gpu_fused.py -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_fused.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_fused.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gen2.json 2> gpurun_out/bench_gen2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_gen2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_gen2.json'))
print(d['value'], d['ms_per_step'], d['phase_ms'], d['roofline']['frac'], {k: v['avg_ms_per_launch'] for k, v in d['roofline'].get('kernels', {}).items()})
PY
