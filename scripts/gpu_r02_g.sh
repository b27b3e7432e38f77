#!/bin/bash
# round 2, GPU call G: after the rank-order FPS kernel and the dense SA tiles -- op-level bench, encoder profile, headline bench
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 300 python scripts/bench_ops.py > gpurun_out/r02_ops_bench.json 2> gpurun_out/r02_ops_bench.err; echo "ops bench rc=$? ${SECONDS}s"
SECONDS=0
timeout 300 python scripts/profile_encode.py > gpurun_out/r02_profile_encode.txt 2>&1; echo "profile rc=$? ${SECONDS}s"
grep -E "^===|Self CUDA time total" gpurun_out/r02_profile_encode.txt
SECONDS=0
python bench.py --steps 5 > gpurun_out/r02_bench_g.json 2> gpurun_out/r02_bench_g.err; echo "bench rc=$? ${SECONDS}s"
python - <<'P'
import json
d = json.load(open("gpurun_out/r02_bench_g.json"))
print(d["value"], d["ms_per_step"], d["phase_ms"], d["roofline"]["frac"], {k: round(v["avg_ms_per_launch"], 4) for k, v in d["roofline"]["kernels"].items()})
print(d["parity"]); print(d.get("fast_mode")); print(d["e2e"]); print(d["encoder"]); print(d["clocks"])
P
