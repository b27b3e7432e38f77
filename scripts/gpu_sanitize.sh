#!/bin/bash
# compute-sanitizer passes over the fused matcher, the front-end and the new ops (small shapes)
set -u
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import helpers, numpy as np
from oracle import reid_oracle as O
from test_frontend_oracle import scene
from pcreid_b200.models.frontend import crop_center_resample
from pcreid_b200.ops import three_nn, three_interpolate
m, orc = helpers.build_pair("pt", (160, 80, 40), device="cuda")
m.set_mode('fast')
t, d = O.synth_objects(3, 160, 0).cuda(), O.synth_objects(4, 160, 1).cuda()
xt, ht = m.encode(t); xd, hd = m.encode(d)
L = m.match_all_pairs(ht, xt, hd, xd)
m2, _ = helpers.build_pair("pt", (256, 128, 64), device="cuda")
m2.set_mode('fast')
t, d = O.synth_objects(5, 256, 0).cuda(), O.synth_objects(7, 256, 1).cuda()
xt, ht = m2.encode(t); xd, hd = m2.encode(d)
L2 = m2.match_all_pairs(ht, xt, hd, xd)
pts, boxes = scene(5000, 7, 3)
out, ln = crop_center_resample(torch.from_numpy(boxes).cuda(), torch.from_numpy(pts).cuda(), 64)
dd, ii = three_nn(t, d[:5, :100].contiguous())
torch.cuda.synchronize()
print("ok", float(L.abs().sum()), float(L2.abs().sum()), int(ln.sum()))
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --kernel-name-exclude kns=at,kns=cub python /tmp/san_case.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|ok " gpurun_out/sanitizer_$tool.log | sort | uniq -c | head -12
done
