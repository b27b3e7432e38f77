#!/bin/bash
# round 2: compute-sanitizer passes over the kernels new in this round (TMA-staged GEMM incl. the 3 x tf32 variant and object maps,
# thread-per-query ball query, fused query_group, fp16 matcher) on small shapes
set -u
mkdir -p gpurun_out
cat > /tmp/san_case2.py <<'PY'
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import helpers
from oracle import reid_oracle as O
import pcreid_b200.kernels as K
from pcreid_b200.ops import ball_query, QueryAndGroup
g = torch.Generator().manual_seed(0)
x1, x2, res = torch.randn(4, 64, 200, generator=g).cuda(), torch.randn(5, 40, 200, generator=g).cuda(), torch.randn(3, 160, 200, generator=g).cuda()
w1, w2 = (torch.randn(64, 160, generator=g) / 8).cuda(), (torch.randn(40, 160, generator=g) / 6).cuda()
m1 = torch.tensor([3, 0, 2, 2, 1, 0], dtype=torch.int32).cuda(); m2 = torch.tensor([4, 4, 0, 1, 3, 2], dtype=torch.int32).cuda()
mr = torch.tensor([0, 2, 1, 1, 2, 0], dtype=torch.int32).cuda()
K._TC_LINEAR["tma"] = True
with K.tensor_core_linear(True):
    a = K.cn_linear(x1, w1, x2=x2, w2=w2, act=1, res=res, x1_map=m1, x2_map=m2, r_map=mr, B=6)
    b = K.cn_linear(torch.randn(3, 512, 132, generator=g).cuda(), (torch.randn(512, 320, generator=g) / 22).cuda(), act=2)     # 256-wide tiles
K._TC_LINEAR.pop("tma")
with K.tensor_core_linear(True, min_k=1 << 30, x3=True):
    c = K.cn_linear(x1, w1, x2=x2, w2=w2, act=1, res=res, x1_map=m1, x2_map=m2, r_map=mr, B=6)
xyz = O.synth_objects(3, 256, 0).cuda()
idx = ball_query(0.0, 0.4, 16, xyz, xyz[:, :128].contiguous())
qg = QueryAndGroup(0.4, 16, use_xyz=True)(xyz, xyz[:, :64].contiguous(), torch.randn(3, 8, 256, generator=g).cuda())
for kind, mode in (("pt", "parity_tc"), ("dgcnn", "fast"), ("pt7m", "fast")):
    m, _ = helpers.build_pair(kind, (128, 64, 32), device="cuda")
    m.set_mode(mode)
    t, d = O.synth_objects(3, 128, 0).cuda(), O.synth_objects(4, 128, 1).cuda()
    xt, ht = m.encode(t); xd, hd = m.encode(d)
    L = m.match_all_pairs(ht, xt, hd, xd)
torch.cuda.synchronize()
print("ok", float(a.abs().sum()), float(b.abs().sum()), float(c.abs().sum()), int(idx.sum()), float(qg.abs().sum()), float(L.abs().sum()))
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --kernel-name-exclude kns=at,kns=cub python /tmp/san_case2.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|ok " gpurun_out/r02_sanitizer_$tool.log | sort | uniq -c | head -12
done
