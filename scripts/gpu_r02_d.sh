#!/bin/bash
# round 2, GPU call D: full GPU suite, default bench (parity_tc headline + fast + strong + parity block), ncu launch list + full capture
set -u
mkdir -p gpurun_out
SECONDS=0
python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^DEBUG\|^INFO" | tail -5; echo "pytest ${SECONDS}s"
SECONDS=0
python bench.py > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; echo "bench rc=$? ${SECONDS}s"
python - <<'P'
import json
d = json.load(open("gpurun_out/r02_bench_d.json"))
print(d["value"], d["ms_per_step"], d["phase_ms"], d["roofline"]["frac"], {k: round(v["avg_ms_per_launch"], 4) for k, v in d["roofline"]["kernels"].items()})
print(d["parity"]); print(d.get("fast_mode")); print(d.get("strong")); print(d["e2e"]); print(d["encoder"])
P
