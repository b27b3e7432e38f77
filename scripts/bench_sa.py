import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pcreid_b200 import kernels as K
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
def timeit(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for C, N, S, k in ((32, 256, 256, 32), (64, 256, 128, 48), (128, 128, 64, 48)):
    g = torch.Generator().manual_seed(0)
    p1 = torch.randn(B, N, C, generator=g).cuda(); cc = torch.randn(B, S, C, generator=g).cuda()
    idx = torch.randint(0, N, (B, S, k), generator=g, dtype=torch.int32).cuda()
    w2 = K.tf32_image(torch.randn(C, C, generator=g) / C ** 0.5).cuda(); w3 = K.tf32_image(torch.randn(C, C, generator=g) / C ** 0.5).cuda()
    b2 = torch.randn(C, generator=g).cuda() * 0.1; b3 = torch.randn(C, generator=g).cuda() * 0.1
    o1 = K.sa_edge_mlp_tc(p1, cc, idx, w2, b2, w3, b3, gen=1); o2 = K.sa_edge_mlp_tc(p1, cc, idx, w2, b2, w3, b3, gen=2)
    t1 = timeit(lambda: K.sa_edge_mlp_tc(p1, cc, idx, w2, b2, w3, b3, gen=1)); t2 = timeit(lambda: K.sa_edge_mlp_tc(p1, cc, idx, w2, b2, w3, b3, gen=2))
    fl = 4.0 * C * C * S * k * B
    print(f"C={C} N={N} S={S} k={k} B={B}: gen1 {t1:.3f} ms  gen2 {t2:.3f} ms ({fl / t2 / 1e9:.1f} TFLOP/s)  max|gen1-gen2| {(o1 - o2).abs().max().item():.2e}", flush=True)
