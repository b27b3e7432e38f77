"""FFMA vs tensor-core cn_linear across shapes (decides the dispatch threshold in kernels.cn_linear)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pcreid_b200.kernels as K
dev = "cuda"
def timeit(fn, iters=5):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
print("B N K CO | ffma_ms tc1_ms tc2_ms | ffma tc1 tc2 TFLOP/s")
for (B, N, Kd, CO) in [(2048,256,32,32),(2048,256,64,64),(2048,256,64,192),(2048,256,128,128),(2048,256,256,256),(1024,256,512,1024),(1024,256,1024,512),(1024,256,512,128),(512,1024,128,128)]:
    x = torch.randn(B, Kd, N, device=dev); w = torch.randn(Kd, CO, device=dev) / Kd ** 0.5
    out = torch.empty(B, CO, N, device=dev)
    with K.tensor_core_linear(False): t0 = timeit(lambda: K.cn_linear(x, w, act=1, out=out))
    K._TC_LINEAR["min_k"] = 8
    K._TC_LINEAR["gen"] = 1
    with K.tensor_core_linear(True): t1 = timeit(lambda: K.cn_linear(x, w, act=1, out=out))
    K._TC_LINEAR["gen"] = 2
    with K.tensor_core_linear(True): t2 = timeit(lambda: K.cn_linear(x, w, act=1, out=out))
    fl = 2.0 * B * N * Kd * CO
    print(B, N, Kd, CO, "|", round(t0, 3), round(t1, 3), round(t2, 3), "|", round(fl / t0 / 1e9, 1), round(fl / t1 / 1e9, 1), round(fl / t2 / 1e9, 1), flush=True)
