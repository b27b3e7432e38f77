"""Timing of the crop -> centre -> resample front-end (csrc/frontend.cu) on one B200: a 250k-point sweep, 300 boxes, 256
samples per box; the reference's points_in_boxes kernel (oracle/_ref/libref_pib.so) timed beside the mask pass."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import frontend_oracle as F
from test_frontend_oracle import scene
from pcreid_b200.models.frontend import crop_center_resample, points_in_boxes_mask
dev = "cuda"
P, B, N = 250000, 300, 256
pts, boxes = scene(P, B, 6)
bt, pt = torch.from_numpy(boxes).to(dev), torch.from_numpy(pts).to(dev)


def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


r = {"P": P, "B": B, "N": N, "mask_pass_ms": timeit(lambda: points_in_boxes_mask(bt, pt)),
     "crop_center_resample_ms": timeit(lambda: crop_center_resample(bt, pt, N)),
     "in_box_tests_per_s": None, "reference_points_in_boxes_kernel": "unavailable"}
r["in_box_tests_per_s"] = P * B / (r["mask_pass_ms"] * 1e-3)
if F.ref_pib_available():
    bl = torch.stack([bt[:, 1], -bt[:, 0], bt[:, 2] + bt[:, 5] * -0.5, bt[:, 4], bt[:, 3], bt[:, 5], bt[:, 6]], 1).unsqueeze(0).contiguous()
    pl = torch.stack([pt[:, 1], -pt[:, 0], pt[:, 2]], 1).unsqueeze(0).contiguous()
    F.ref_points_in_boxes_lidar(bl, pl)
    t0 = time.perf_counter()
    for _ in range(5): F.ref_points_in_boxes_lidar(bl, pl)
    r["reference_points_in_boxes_kernel"] = {"ms_incl_300MB_int_mask_zero_fill": (time.perf_counter() - t0) / 5 * 1e3,
                                             "note": "reference writes an int32 (P, B) mask; ours 1 bit per (point, box)"}
cpu0 = time.perf_counter()
rank = torch.randint(0, 50, (B, N)).numpy()
F.crop_center_resample(boxes, pts, N, rank % 1)
r["numpy_restatement_cpu_ms"] = (time.perf_counter() - cpu0) * 1e3
print(json.dumps(r))
