import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, helpers
from torch.profiler import profile, ProfilerActivity
from oracle import reid_oracle as O
dev = "cuda"
m, _ = helpers.build_pair("pt", (256, 128, 64), device=dev, perturb=False)
m.set_mode(sys.argv[1] if len(sys.argv) > 1 else 'fast')
x = O.synth_objects(2048, 256, 0).to(dev)
for _ in range(3): m.encode(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m.encode(x); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
