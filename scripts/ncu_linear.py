"""single-shape driver for `ncu --set full` on the TMA-staged cn_linear kernel: python scripts/ncu_linear.py B N K CO"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pcreid_b200.kernels as K  # noqa: E402

B, N, Kd, CO = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (2048, 256, 128, 128)
x = torch.randn(B, Kd, N, device="cuda"); w = torch.randn(Kd, CO, device="cuda") / Kd ** 0.5
out = torch.empty(B, CO, N, device="cuda")
K._TC_LINEAR["tma"] = True
with K.tensor_core_linear(True):
    for _ in range(3):
        K.cn_linear(x, w, act=1, out=out)
torch.cuda.synchronize()
