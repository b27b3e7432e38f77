"""Per-kernel CUDA time of one 1024 x 1024 match in a tensor-core mode (torch profiler): what is left beside the three fused kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from pcreid_b200 import synthetic as S
from pcreid_b200.models import build_model
dev = "cuda"
mode = sys.argv[1] if len(sys.argv) > 1 else "parity_tc"
torch.manual_seed(66)
m = build_model(S.point_transformer_cfg((256, 128, 64))).eval().to(dev)
m.set_mode(mode)
t, d = S.synth_objects(1024, 256, 1000).to(dev), S.synth_objects(1024, 256, 1).to(dev)
xt, ht = m.encode(t); xd, hd = m.encode(d)
for _ in range(2): m.match_all_pairs(ht, xt, hd, xd)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    m.match_all_pairs(ht, xt, hd, xd); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
