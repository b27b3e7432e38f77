import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pcreid_b200 import kernels as K
B = 2048
C, N, S, k = {32: (32, 256, 256, 32), 64: (64, 256, 128, 48), 128: (128, 128, 64, 48)}[int(sys.argv[1])]
g = torch.Generator().manual_seed(0)
p1 = torch.randn(B, N, C, generator=g).cuda(); cc = torch.randn(B, S, C, generator=g).cuda()
idx = torch.randint(0, N, (B, S, k), generator=g, dtype=torch.int32).cuda()
w2 = K.tf32_image(torch.randn(C, C, generator=g) / C ** 0.5).cuda(); w3 = K.tf32_image(torch.randn(C, C, generator=g) / C ** 0.5).cuda()
b2 = torch.randn(C, generator=g).cuda() * 0.1; b3 = torch.randn(C, generator=g).cuda() * 0.1
for _ in range(2): K.sa_edge_mlp_tc(p1, cc, idx, w2, b2, w3, b3, gen=2)
torch.cuda.synchronize()
