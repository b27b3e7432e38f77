"""per-kernel CUDA time of the fused attention blocks: python scripts/profile_attn.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from torch.profiler import profile, ProfilerActivity
from pcreid_b200.models.pointnet2_utils import Self_Attention, FP_SA
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
scale = torch.tensor([2.0, 0.9, 0.8], device=dev)
for d, S in ((32, 256), (64, 128), (128, 64)):
    m = Self_Attention(d, 2).to(dev).eval(); m.tc_mode = True
    feat, xyz = torch.randn(B, d, S, device=dev), torch.randn(B, S, 3, device=dev) * scale
    with torch.no_grad():
        for _ in range(3): m(feat, xyz)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            m(feat, xyz); torch.cuda.synchronize()
    print(f"== SA d={d} S={S}")
    for e in prof.key_averages():
        print(f"   {e.key[:70]:70s} {e.device_time_total:9.1f} us  x{e.count}")
for f1, f2, d, out, N, S, pm in ((64, 128, 64, 128, 128, 64, False), (32, 128, 64, 64, 256, 128, False), (3, 64, 64, 32, 256, 256, True)):
    m = FP_SA(0, f1, f2, d, out, 2).to(dev).eval(); m.tc_mode = True
    xyz1, xyz2 = torch.randn(B, N, 3, device=dev) * scale, torch.randn(B, S, 3, device=dev) * scale
    feat1 = xyz1.contiguous() if pm else torch.randn(B, f1, N, device=dev)
    feat2 = torch.randn(B, f2, S, device=dev)
    with torch.no_grad():
        for _ in range(3): m(feat1, xyz1, feat2, xyz2, feat1_point_major=pm)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            m(feat1, xyz1, feat2, xyz2, feat1_point_major=pm); torch.cuda.synchronize()
    print(f"== FP {f1, f2, d, out, N, S}")
    for e in prof.key_averages():
        print(f"   {e.key[:70]:70s} {e.device_time_total:9.1f} us  x{e.count}")
