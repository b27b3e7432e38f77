#!/bin/bash
# A/B of the library at HEAD (working tree) against point-cloud-reid_b200/build/libprev.so (a build of an earlier commit):
# fused-matcher tests at HEAD, a 96 x 80 logit dump of each (compared bit for bit), kernel times from bench.py, alternating
set -u
mkdir -p gpurun_out
cat > /tmp/dump_logits.py <<'PY'
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import helpers
from oracle import reid_oracle as O
m, _ = helpers.build_pair("pt", (256, 128, 64), device="cuda")
m.set_mode(sys.argv[3] if len(sys.argv) > 3 else "parity_tc")
t, d = O.synth_objects(96, 256, 0).cuda(), O.synth_objects(80, 256, 1).cuda()
xt, ht = m.encode(t); xd, hd = m.encode(d)
L = m.match_all_pairs(ht, xt, hd, xd).cpu()
torch.save(L, sys.argv[1])
if len(sys.argv) > 2 and sys.argv[2] != "-":
    R = torch.load(sys.argv[2])
    print("bit-identical to the previous build:", bool(torch.equal(L, R)), "max abs diff", float((L - R).abs().max()))
PY
LIB=point-cloud-reid_b200/libpcreid_sm100.so
cp $LIB /tmp/libhead.so
timeout 300 python -m pytest tests/test_gpu_fused.py tests/test_gpu_image.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -1
run() {  # $1 = tag
  timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extra > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err
  python - <<P
import json
d = json.load(open("gpurun_out/ab_$1.json"))
print("$1:", round(d["value"]), round(d["ms_per_step"], 2), {k: round(x["avg_ms_per_launch"], 4) for k, x in d["roofline"]["kernels"].items()}, d["clocks"]["sm_mhz"])
P
}
for mode in parity_tc fast; do
  cp point-cloud-reid_b200/build/libprev.so $LIB; PCREID_B7_BYTES=${PREV_B7_BYTES:-10240} timeout 200 python /tmp/dump_logits.py /tmp/logits_prev_$mode.pt - $mode
  cp /tmp/libhead.so $LIB; timeout 200 python /tmp/dump_logits.py /tmp/logits_head_$mode.pt /tmp/logits_prev_$mode.pt $mode
done
for i in 1 2; do
  cp point-cloud-reid_b200/build/libprev.so $LIB; PCREID_B7_BYTES=${PREV_B7_BYTES:-10240} run prev$i
  cp /tmp/libhead.so $LIB; run head$i
done
