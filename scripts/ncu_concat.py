"""one all-pairs 'concat' match (4096 x 4096, tensor-core head) for an ncu capture of pair_concat_head_tc_kernel"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, helpers
m, _ = helpers.build_pair("concat", (128, 64, 32), device="cuda", perturb=False)
m.set_mode('fast')
T = D = 4096
e_t, e_d = torch.randn(T, 128, device="cuda"), torch.randn(D, 128, device="cuda")
with torch.no_grad():
    for _ in range(2):
        L = m.concat_all_pairs_pooled(e_t, e_d)
torch.cuda.synchronize()
print(float(L.abs().mean()))
