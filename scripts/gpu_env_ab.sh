#!/bin/bash
# A/B of an environment knob of the fused matcher kernels: usage  gpu_env_ab.sh VAR v1 v2 ...
# per value: fused-matcher tests, a 96 x 80 logit dump (compared bit for bit with the first value's), kernel times from bench.py
set -u
mkdir -p gpurun_out
VAR=$1; shift
cat > /tmp/dump_logits.py <<'PY'
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import helpers
from oracle import reid_oracle as O
m, _ = helpers.build_pair("pt", (256, 128, 64), device="cuda")
m.set_mode("parity_tc")
t, d = O.synth_objects(96, 256, 0).cuda(), O.synth_objects(80, 256, 1).cuda()
xt, ht = m.encode(t); xd, hd = m.encode(d)
L = m.match_all_pairs(ht, xt, hd, xd).cpu()
torch.save(L, sys.argv[1])
if len(sys.argv) > 2:
    R = torch.load(sys.argv[2])
    print("bit-identical to first:", bool(torch.equal(L, R)), "max abs diff", float((L - R).abs().max()))
PY
first=""
for v in "$@"; do
  export $VAR=$v
  timeout 300 python -m pytest tests/test_gpu_fused.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -1
  timeout 200 python /tmp/dump_logits.py /tmp/logits_$v.pt $first
  [ -z "$first" ] && first=/tmp/logits_$v.pt
  timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extra > gpurun_out/ab_${VAR}_$v.json 2> gpurun_out/ab_${VAR}_$v.err
  python - <<P
import json
d = json.load(open("gpurun_out/ab_${VAR}_$v.json"))
print("$VAR=$v:", round(d["value"]), round(d["ms_per_step"], 2), {k: round(x["avg_ms_per_launch"], 4) for k, x in d["roofline"]["kernels"].items()}, d["clocks"]["sm_mhz"])
P
done
