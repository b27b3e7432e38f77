#!/bin/bash
# round 2, GPU call C: fused-matcher tests + torch-op tests + short bench
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused.py tests/test_gpu_image.py tests/test_gpu_torch_ops.py tests/test_gpu_model.py -x -q 2>&1 | grep -v "^DEBUG\|^INFO" | tail -15
SECONDS=0
python bench.py --no-extra > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; echo "bench rc=$? ${SECONDS}s"
python - <<'P'
import json
d = json.load(open("gpurun_out/r02_bench_c.json"))
print(d["value"], d["ms_per_step"], d["phase_ms"], d["roofline"]["frac"], {k: round(v["avg_ms_per_launch"], 4) for k, v in d["roofline"]["kernels"].items()})
print(d["parity"])
P
