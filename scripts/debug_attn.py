"""Fused attention blocks vs the unfused chain: error + time per block (B200).  python scripts/debug_attn.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from pcreid_b200.models.pointnet2_utils import Self_Attention, FP_SA
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048


def timeit(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


scale = torch.tensor([2.0, 0.9, 0.8], device=dev)
for d, S in ((32, 256), (64, 128), (128, 64)):
    torch.manual_seed(0)
    m = Self_Attention(d, 2).to(dev).eval()
    feat, xyz = torch.randn(B, d, S, device=dev), torch.randn(B, S, 3, device=dev) * scale
    with torch.no_grad():
        ref = m(feat, xyz); t0 = timeit(lambda: m(feat, xyz))
        m.tc_mode = True
        got = m(feat, xyz); t1 = timeit(lambda: m(feat, xyz))
    print(f"SA d={d} S={S}: err {(got - ref).abs().max().item():.2e} nan {int(torch.isnan(got).sum())}  unfused {t0:.3f} ms  fused {t1:.3f} ms", flush=True)
for f1, f2, d, out, N, S, pm in ((64, 128, 64, 128, 128, 64, False), (32, 128, 64, 64, 256, 128, False), (3, 64, 64, 32, 256, 256, True)):
    torch.manual_seed(0)
    m = FP_SA(0, f1, f2, d, out, 2).to(dev).eval()
    xyz1, xyz2 = torch.randn(B, N, 3, device=dev) * scale, torch.randn(B, S, 3, device=dev) * scale
    feat1 = xyz1.contiguous() if pm else torch.randn(B, f1, N, device=dev)
    feat2 = torch.randn(B, f2, S, device=dev)
    with torch.no_grad():
        ref = m(feat1, xyz1, feat2, xyz2, feat1_point_major=pm); t0 = timeit(lambda: m(feat1, xyz1, feat2, xyz2, feat1_point_major=pm))
        m.tc_mode = True
        got = m(feat1, xyz1, feat2, xyz2, feat1_point_major=pm); t1 = timeit(lambda: m(feat1, xyz1, feat2, xyz2, feat1_point_major=pm))
    print(f"FP {f1,f2,d,out,N,S}: err {(got - ref).abs().max().item():.2e} nan {int(torch.isnan(got).sum())}  unfused {t0:.3f} ms  fused {t1:.3f} ms", flush=True)
