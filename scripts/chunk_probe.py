"""Fused matcher vs pair-chunk size: does keeping a chunk's spilled stage-1 images inside the 126 MB L2 pay?
Sums the per-launch CUDA-event time of the three fused kernels over one 1024 x 1024 match for several chunk sizes."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcreid_b200 import synthetic as S  # noqa: E402
from pcreid_b200.models import build_model  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(66)
model = build_model(S.point_transformer_cfg((256, 128, 64))).eval().to(dev)
model.set_mode("parity_tc")
T = D = 1024
t, d = S.synth_objects(T, 256, 1000).to(dev), S.synth_objects(D, 256, 1).to(dev)
xt, ht = model.encode(t)
xd, hd = model.encode(d)
fm = model.fused_matcher()
ref = None
for chunk in [int(c) for c in os.environ.get("CHUNKS", "65536,16384,4096,2048,1024,512").split(",")]:
    import pcreid_b200.models.ReIDNet as R
    fn = lambda: model.match_all_pairs(ht, xt, hd, xd, chunk=chunk, _exact_chunk=True)
    out = fn()
    torch.cuda.synchronize()
    if ref is None:
        ref = out
    same = bool(torch.equal(out, ref))
    fm.timing = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    k = {}
    for name, a0, a1, units in fm.timing:
        k[name] = k.get(name, 0.0) + a0.elapsed_time(a1)
    fm.timing = None
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    fn()
    e3.record()
    torch.cuda.synchronize()
    print(json.dumps({"chunk_pairs": chunk, "scratch_MB": chunk * (2 * 2 * 16384 + 2 * 10240 + 1024) / 1e6, "kernels_ms": k,
                      "kernels_sum_ms": sum(k.values()), "wall_ms_with_events": e0.elapsed_time(e1), "wall_ms": e2.elapsed_time(e3),
                      "bit_identical_to_first": same}), flush=True)
