"""one fused attention block for ncu: python scripts/ncu_attn.py sa|fp d"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pcreid_b200.models.pointnet2_utils import Self_Attention, FP_SA
dev = "cuda"
B = 2048
scale = torch.tensor([2.0, 0.9, 0.8], device=dev)
kind = sys.argv[1]
if kind == "sa":
    d = int(sys.argv[2]); S = 8192 // d
    m = Self_Attention(d, 2).to(dev).eval(); m.tc_mode = True
    feat, xyz = torch.randn(B, d, S, device=dev), torch.randn(B, S, 3, device=dev) * scale
    with torch.no_grad():
        for _ in range(2): m(feat, xyz)
else:
    f1, f2, d, out, N, S = 32, 128, 64, 64, 256, 128
    m = FP_SA(0, f1, f2, d, out, 2).to(dev).eval(); m.tc_mode = True
    xyz1, xyz2 = torch.randn(B, N, 3, device=dev) * scale, torch.randn(B, S, 3, device=dev) * scale
    feat1, feat2 = torch.randn(B, f1, N, device=dev), torch.randn(B, f2, S, device=dev)
    with torch.no_grad():
        for _ in range(2): m(feat1, xyz1, feat2, xyz2)
torch.cuda.synchronize()
