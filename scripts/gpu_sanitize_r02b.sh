#!/bin/bash
# round 2, second sanitizer pass: the kernels added late in the round -- fps_rank_kernel (warp-per-object, multi-warp, padded slot
# counts, torch-path mode, temp write-back) and the dense-tile sa_edge_mlp_tc2_kernel (straddling centres, ragged last tile, units
# that split an object) -- memcheck / racecheck / synccheck on small shapes
set -u
mkdir -p gpurun_out
cat > /tmp/san_case3.py <<'PY'
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import reid_oracle as O
import pcreid_b200.kernels as K
from pcreid_b200.ops import furthest_point_sample
g = torch.Generator().manual_seed(0)
acc = 0
for B, N, M in ((300, 256, 16), (3, 1000, 12), (2, 3000, 8), (2, 4096, 8)):
    x = O.synth_objects(B, N, 1, dup=True).cuda()
    acc += int(furthest_point_sample(x, M).sum())
x = O.synth_objects(4, 200, 2).cuda()
acc += int(K.farthest_point_sample(x, 24, start=torch.tensor([3, 0, 199, 17])).sum())
tot = 0.0
for C, N, S, k, B in ((32, 96, 40, 32, 3), (64, 128, 50, 48, 2), (128, 64, 21, 48, 5), (64, 64, 33, 16, 2), (32, 128, 50, 63, 2), (128, 128, 64, 48, 300)):
    p1, cc = torch.randn(B, N, C, generator=g).cuda(), torch.randn(B, S, C, generator=g).cuda()
    idx = torch.randint(0, N, (B, S, k), generator=g, dtype=torch.int32).cuda()
    w2, w3 = K.tf32_image(torch.randn(C, C, generator=g) / C ** 0.5).cuda(), K.tf32_image(torch.randn(C, C, generator=g) / C ** 0.5).cuda()
    b2, b3 = (torch.randn(C, generator=g) * 0.1).cuda(), (torch.randn(C, generator=g) * 0.1).cuda()
    tot += float(K.sa_edge_mlp_tc(p1, cc, idx, w2, b2, w3, b3, gen=2).abs().sum())
torch.cuda.synchronize()
print("ok", acc, round(tot, 2))
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --kernel-name-exclude kns=at,kns=cub python /tmp/san_case3.py > gpurun_out/r02b_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|ok " gpurun_out/r02b_sanitizer_$tool.log | sort | uniq -c | head -12
done
