"""Op-level benchmark of the five mmdet3d point ops + the torch-path kNNs on one B200: CUDA-event timing, algorithmic
bytes / time against the measured HBM peak, distance evaluations per second, and the reference's own
.cu kernels (oracle/_ref, compiled unmodified for sm_100a) timed beside ours as the like-for-like bar.
Writes one JSON document to stdout."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from oracle import ops_oracle as P
from oracle import reid_oracle as O
import pcreid_b200.kernels as K
from pcreid_b200.ops import (ball_query, furthest_point_sample, gather_points, grouping_operation, knn, three_interpolate,
                             three_nn)

dev = "cuda"
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ALU_EVALS_PER_S = 148 * 128 * 1.965e9 / 6.0      # distance evaluations / s if the fp32 pipes did nothing else (6.2e12)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()                      # L2 flush between timed iterations (buffer > 126 MB L2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


rows = []


def add(name, ms, bytes_algo, evals=None, ref_ms=None, shape=""):
    r = {"op": name, "shape": shape, "ms": ms, "algorithmic_GB": bytes_algo / 1e9, "achieved_GBps": bytes_algo / ms / 1e6,
         "frac_of_measured_hbm": bytes_algo / ms / 1e6 / peak}
    if evals:
        r["distance_evals_per_s"] = evals / (ms * 1e-3)
        # ALU roofline of the distance arithmetic alone: 6 fp32-pipe instructions per evaluation (3 subtractions, 1 multiply,
        # 2 fused multiply-adds; selection / bookkeeping not counted) on 148 SMs x 128 fp32 lanes at the maximum SM clock
        r["frac_of_fp32_alu_roofline"] = r["distance_evals_per_s"] / ALU_EVALS_PER_S
    if ref_ms:
        r["reference_cu_ms"] = ref_ms
        r["speedup_vs_reference_cu"] = ref_ms / ms
    rows.append(r)


have_ref = P.ref_available()
B = 2048
for (N, S, k) in ((256, 256, 32), (256, 128, 48), (1024, 512, 48)):
    x = O.synth_objects(B, N, 0).to(dev)
    c = x[:, :S].contiguous()
    ms = timeit(lambda: knn(k, x, c))
    ref = timeit(lambda: P.ref_knn(k, x, c)) if have_ref else None
    add("knn (mmdet3d op)", ms, B * (12 * (N + S) + 8 * S * k), B * S * N, ref, f"B={B} N={N} S={S} k={k}")
    ms = timeit(lambda: K.knn_point(k, x, c))
    add("knn_point (torch-path arithmetic)", ms, B * (12 * (N + S) + 4 * S * k), B * S * N, None, f"B={B} N={N} S={S} k={k}")
for (N, M) in ((256, 128), (1024, 256), (4096, 512)):
    b = B if N <= 1024 else 64
    x = O.synth_objects(b, N, 1).to(dev)
    ms = timeit(lambda: furthest_point_sample(x, M))
    ref = timeit(lambda: P.ref_furthest_point_sample(x, M)) if have_ref else None
    add("furthest_point_sample", ms, b * (12 * N + 4 * M), b * M * N, ref, f"B={b} N={N} M={M}")
x = O.synth_objects(B, 256, 2).to(dev)
c = x[:, :128].contiguous()
ms = timeit(lambda: ball_query(0.0, 0.8, 32, x, c))
ref = timeit(lambda: P.ref_ball_query(0.0, 0.8, 32, x, c)) if have_ref else None
add("ball_query", ms, B * (12 * (256 + 128) + 4 * 128 * 32), B * 128 * 256, ref, f"B={B} N=256 S=128 k=32")
f = torch.randn(B, 64, 256, device=dev)
idx = torch.randint(0, 256, (B, 128, 48), device=dev, dtype=torch.int32)
ms = timeit(lambda: grouping_operation(f, idx))
ref = timeit(lambda: P.ref_grouping_operation(f, idx)) if have_ref else None
add("grouping_operation", ms, B * (4 * 128 * 48 + 4 * 64 * 256 + 4 * 64 * 128 * 48), None, ref, f"B={B} C=64 N=256 S=128 k=48")
i2 = torch.randint(0, 256, (B, 128), device=dev, dtype=torch.int32)
ms = timeit(lambda: gather_points(f, i2))
ref = timeit(lambda: P.ref_gather_points(f, i2)) if have_ref else None
add("gather_points", ms, B * (4 * 128 + 4 * 64 * 256 + 4 * 64 * 128), None, ref, f"B={B} C=64 N=256 M=128")
# feature propagation ops: 256 fine points interpolate from 64 coarse points, 128 channels
tgt, srcp = O.synth_objects(B, 256, 3).to(dev), O.synth_objects(B, 64, 4).to(dev)
ms = timeit(lambda: three_nn(tgt, srcp))
ref = timeit(lambda: P.ref_three_nn(tgt, srcp)) if have_ref else None
add("three_nn", ms, B * (12 * (256 + 64) + 24 * 256), B * 256 * 64, ref, f"B={B} N=256 M=64")
fz = torch.randn(B, 128, 64, device=dev)
dz, iz = three_nn(tgt, srcp)
wz = torch.softmax(-dz, 2).contiguous()
ms = timeit(lambda: three_interpolate(fz, iz, wz))
ref = timeit(lambda: P.ref_three_interpolate(fz, iz, wz)) if have_ref else None
add("three_interpolate", ms, B * (24 * 256 + 4 * 128 * 64 + 4 * 128 * 256), None, ref, f"B={B} C=128 M=64 N=256")
xf = torch.randn(B, 64, 256, device=dev)
ms = timeit(lambda: K.knn_feature(xf, 20))
add("knn_feature (DGCNN)", ms, B * (4 * 64 * 256 + 4 * 256 * 20), B * 256 * 256, None, f"B={B} C=64 N=256 k=20")
print(json.dumps({"hbm_peak_GBps": peak, "peak_source": "MEASURED_PEAKS.json", "l2_flush": "256 MB zero-fill between iterations",
                  "alu_roofline": {"distance_evals_per_s": ALU_EVALS_PER_S, "how": "6 fp32-pipe instructions per distance evaluation (3 sub, "
                                   "1 mul, 2 fma; selection / bookkeeping not counted) x 148 SMs x 128 fp32 lanes x 1.965 GHz"},
                  "reference_cu": "oracle/_ref (reference .cu unmodified, sm_100a)" if have_ref else "unavailable", "rows": rows}, indent=1))
