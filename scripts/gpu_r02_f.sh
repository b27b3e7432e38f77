#!/bin/bash
# round 2, GPU call F: full GPU suite with the TMA-staged GEMM wired into the models, C3 (DGCNN) in parity_tc and fast mode, headline bench
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^DEBUG\|^INFO" | tail -12; echo "pytest ${SECONDS}s"
for mode in parity_tc fast; do
SECONDS=0
python bench.py --config c3 --mode $mode --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c3_$mode.json 2> gpurun_out/r02_bench_c3_$mode.err; echo "c3 $mode rc=$? ${SECONDS}s"
python - <<P
import json
d = json.load(open("gpurun_out/r02_bench_c3_$mode.json"))
print(d["value"], d["ms_per_step"], d["phase_ms"], d["encoder"])
P
done
SECONDS=0
python bench.py --steps 5 > gpurun_out/r02_bench_f.json 2> gpurun_out/r02_bench_f.err; echo "bench rc=$? ${SECONDS}s"
python - <<'P'
import json
d = json.load(open("gpurun_out/r02_bench_f.json"))
print(d["value"], d["ms_per_step"], d["phase_ms"], d["roofline"]["frac"], {k: round(v["avg_ms_per_launch"], 4) for k, v in d["roofline"]["kernels"].items()})
print(d["parity"]); print(d.get("fast_mode")); print(d["e2e"]); print(d["encoder"])
P
