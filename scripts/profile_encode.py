"""Per-kernel CUDA time of one encoder pass (torch profiler): python scripts/profile_encode.py [mode] [objects] [points] [kind]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, helpers
from torch.profiler import profile, ProfilerActivity
from oracle import reid_oracle as O
dev = "cuda"
mode = sys.argv[1] if len(sys.argv) > 1 else 'fast'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
N = int(sys.argv[3]) if len(sys.argv) > 3 else 256
kind = sys.argv[4] if len(sys.argv) > 4 else "pt"
m, _ = helpers.build_pair(kind, (N, N // 2, N // 4), device=dev, perturb=False)
m.set_mode(mode)
x = O.synth_objects(B, N, 0).to(dev)
for _ in range(3): m.encode(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m.encode(x); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
