"""furthest_point_sample at the op-bench shapes (+ the torch-path sampler) against the reference .cu (oracle/_ref), with the
kernel choice taken from the environment (PCREID_FPS_PPT: -1 = register kernel, 0 = heuristic, 4/8/16/32 = slots per thread of
the shared-memory rank-order kernel).  One JSON line per shape; indices are compared with the reference kernel's."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from oracle import ops_oracle as P
from oracle import reid_oracle as O
import pcreid_b200.kernels as K
from pcreid_b200.ops import furthest_point_sample

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=7, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


have_ref = P.ref_available()
knob = os.environ.get("PCREID_FPS_PPT", "0")
shapes = [(2048, 256, 128), (2048, 1024, 256), (2048, 1024, 512), (256, 1024, 256), (64, 4096, 512), (8, 4096, 512), (16, 8192, 1024)]
for (b, N, M) in shapes:
    x = O.synth_objects(b, N, 1).to(dev)
    got = furthest_point_sample(x, M)
    ms = timeit(lambda: furthest_point_sample(x, M))
    row = {"knob": knob, "shape": f"B={b} N={N} M={M}", "ms": round(ms, 4), "distance_evals_per_s": b * M * N / (ms * 1e-3)}
    if have_ref:
        ref = P.ref_furthest_point_sample(x, M)
        row["equal_to_reference_cu"] = bool(torch.equal(got, ref))
        row["reference_cu_ms"] = round(timeit(lambda: P.ref_furthest_point_sample(x, M)), 4)
        row["speedup_vs_reference_cu"] = round(row["reference_cu_ms"] / ms, 3)
    print(json.dumps(row), flush=True)
# torch-path sampler of the ReID SA layers (sampling='FPS')
x = O.synth_objects(2048, 256, 3).to(dev)
start = torch.zeros(2048, dtype=torch.long)
ms = timeit(lambda: K.farthest_point_sample(x, 128, start=start))
print(json.dumps({"knob": knob, "shape": "torch path B=2048 N=256 M=128", "ms": round(ms, 4)}), flush=True)
