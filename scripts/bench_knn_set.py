import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pcreid_b200 import kernels as K
from oracle import reid_oracle as O
B = 2048
def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for N, S, k in ((256, 256, 32), (256, 128, 48), (128, 64, 48), (512, 512, 32), (512, 256, 48), (1024, 1024, 32), (1024, 512, 48)):
    b = B if N <= 256 else 256
    x = O.synth_objects(b, N, 0).cuda().contiguous(); q = x[:, :S].contiguous()
    print(f"N={N} S={S} k={k} B={b}: ordered {timeit(lambda: K.knn_point(k, x, q)):.3f} ms   set {timeit(lambda: K.knn_point_set(k, x, q)):.3f} ms", flush=True)
