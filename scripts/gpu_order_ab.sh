#!/bin/bash
# A/B: unit order handed to phase 1b (runs of equal search object vs runs of equal template); the matrix of
# profiles/r02_unit_order_ab.json also had the resident-search-tile phase-1a kernel (dropped, see pair_tc2.cu)
set -u
mkdir -p gpurun_out
run() {  # $1 = tag
  timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extra > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err
  python - <<P
import json
d = json.load(open("gpurun_out/ab_$1.json"))
print("$1:", round(d["value"]), round(d["ms_per_step"], 2), {k: round(x["avg_ms_per_launch"], 4) for k, x in d["roofline"]["kernels"].items()}, d["clocks"]["sm_mhz"])
P
}
timeout 300 python -m pytest tests/test_gpu_fused.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -1
for i in 1 2; do
  PCREID_P1B_ORDER=templ run bt_$i
  PCREID_P1B_ORDER=search run bs_$i
done
