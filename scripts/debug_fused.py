"""GPU diagnostic for the fused tensor-core matcher: compares every intermediate it exposes with the fp32 parity
kernels (same device) so that a mismatch is localised in one run.  Usage: python scripts/debug_fused.py [N]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import helpers
from oracle import reid_oracle as O
from pcreid_b200 import kernels as K
from pcreid_b200.models import fused_pairs as FP

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = "cuda"
m, orc = helpers.build_pair("pt", (N, N // 2, N // 4), device=dev)
T, D = 3, 4
t, d = O.synth_objects(T, N, 0).to(dev), O.synth_objects(D, N, 1).to(dev)
xt, ht = m.encode(t)
xd, hd = m.encode(d)
X1, X2 = m.cross_stage1, m.cross_stage2
ti = torch.arange(T, device=dev).repeat_interleave(D)
dj = torch.arange(D, device=dev).repeat(T)
ti32, dj32 = ti.int().contiguous(), dj.int().contiguous()


def rep(name, got, ref):
    err = (got.float() - ref.float()).abs().max().item()
    print(f"{name:28s} max|err| {err:.3e}   ref scale {ref.abs().max().item():.3e}", flush=True)


f = FP.FusedXcorr(m)
pt, pd = f.prepare(ht, xt), f.prepare(hd, xd)
pk1 = X1.packed()
q1 = K.cn_linear(ht, pk1["q"])
rep("pack QF1", FP.decode_image(pt.QF1).permute(0, 2, 1, 3).reshape(T, 64, N), torch.nn.functional.elu(q1) + 1)
rep("pack H (+beta2 in gen2)", FP.decode_image(pt.H).permute(0, 2, 1, 3).reshape(T, 64, N), ht + (f._b2_1[None, :, None] if f.gen2 else 0))
dbg = {}
logits = f.match(pt, pd, ti, dj, debug=dbg)
torch.cuda.synchronize()
# parity-path intermediates
NT = N // 128
q_t, q_d = X1.search_query(ht), X1.search_query(hd)
wkv_t, ks_t = X1.template_summary(ht, X1.position_code(xt))
wkv_d, ks_d = X1.template_summary(hd, X1.position_code(xd))
a = X1.attend(ht, q_t, wkv_d, ks_d, N, s_map=ti32, t_map=dj32)
b = X1.attend(hd, q_d, wkv_t, ks_t, N, s_map=dj32, t_map=ti32)
A = FP.decode_image(dbg["A"].view(-1, 2, NT, 8, 128, 16))            # (P, 2, NT, 64, 128)
A = A.permute(0, 1, 3, 2, 4).reshape(-1, 2, 64, N)
rep("stage1 a (role 0)", A[:, 0], a)
rep("stage1 b (role 1)", A[:, 1], b)
pos2_t, pos2_d = X2.position_code(xt), X2.position_code(xd)
wkv_b, ks_b = X2.template_summary(b, pos2_d, pos_map=dj32)
wkv_a, ks_a = X2.template_summary(a, pos2_t, pos_map=ti32)
pk2 = X2.packed()
for role, (wkv, ks) in enumerate(((wkv_a, ks_a), (wkv_b, ks_b))):
    M = K.cn_linear(wkv, pk2["merge"], x1_pm=True, y_pm=True) * N       # (P, d, out)
    if f.gen2:
        M = M - M.mean(2, keepdim=True)                                  # centred merge weight
    B7 = FP.decode_b7(dbg["B7"][:, role])                               # (P, 64, 144)
    exp = torch.zeros_like(B7)
    exp[:, :32, :64] = M[:, :32]
    exp[:, 32:, 64:128] = M[:, 32:]
    exp[:, :32, 128] = ks[:, :32]
    exp[:, 32:, 129] = ks[:, 32:]
    rep(f"B7 role {role}", B7, exp)
o1 = X2.attend(a, X2.search_query(a), wkv_b, ks_b, N)
o2 = X2.attend(b, X2.search_query(b), wkv_a, ks_a, N)
part = dbg["part"]
bb = f._b2_2[None, :] if f.gen2 else 0          # gen2 adds LayerNorm2's beta after the pooling
rep("pool max role0", part[:, 0, :64] + bb, o1.max(2)[0])
rep("pool sum role0", part[:, 0, 64:] + N * bb, o1.sum(2))
rep("pool max role1", part[:, 1, :64] + bb, o2.max(2)[0])
rep("pool sum role1", part[:, 1, 64:] + N * bb, o2.sum(2))
pooled_ref = K.cn_pool(o1, o2, mode=0, transposed=True)
rep("pooled", dbg["pooled"], pooled_ref)
ref_logits = m._head_cn(pooled_ref)
rep("logits vs parity path", logits, ref_logits)
Lo = orc.match_all_pairs(*[x.cpu() for x in (ht, xt, hd, xd)])
rep("logits vs oracle", logits.cpu().view(T, D), Lo)
print("logit std", Lo.std().item(), "top1 agree", (logits.cpu().view(T, D).argmax(1) == Lo.argmax(1)).float().mean().item())
