#!/bin/bash
# A/B of the column-split (two threads per tile row) variants of the fused matcher kernels: tests + kernel times per variant
set -u
mkdir -p gpurun_out
for v in "${@:-1 2}"; do
  export PCREID_P1B_SPLIT=$v
  timeout 300 python -m pytest tests/test_gpu_fused.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -2
  timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extra > gpurun_out/split_ab_$v.json 2> gpurun_out/split_ab_$v.err
  python - <<P
import json
d = json.load(open("gpurun_out/split_ab_$v.json"))
print("split $v:", round(d["value"]), round(d["ms_per_step"], 2), {k: round(x["avg_ms_per_launch"], 4) for k, x in d["roofline"]["kernels"].items()}, d["clocks"]["sm_mhz"])
P
done
