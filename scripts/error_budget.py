#!/usr/bin/env python
"""CPU study: where does the 16-bit-operand error of the fused matcher come from?

A torch emulation of the fused `xcorr_eff` arithmetic (models/fused_pairs.py + csrc/pair_tc*.cu) in which every place where
the kernels round a value to the 16-bit operand format is a named SITE that can be switched on or off.  For each configuration
the logits of a block of pairs are compared with the all-fp32 run: rms / max error and raw top-1 agreement.  Runs on the CPU
with the oracle's weights (seed 66) and fp32 oracle embeddings, i.e. it isolates the matcher.

  python scripts/error_budget.py [--fmt f16|bf16] [--tracks 64] [--dets 128] [--npts 256]
"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SITES = ["QF1", "MK1", "X1", "W1", "Hop", "Hres", "Hd1", "a", "a_res", "PV", "KfV", "KVb", "B7", "W1b", "W2", "Qf2", "X2", "Hd2"]


class Rounder:
    def __init__(self, fmt, on):
        self.dt = torch.float16 if fmt == "f16" else torch.bfloat16
        self.on = set(on)

    def __call__(self, x, site):
        return x.to(self.dt).float() if site in self.on else x


def center(w):
    return w - w.mean(0, keepdim=True)


def elu1(x):
    return F.elu(x) + 1


def ln_noaffine(x, eps=1e-5):
    return x * torch.rsqrt((x * x).mean(-1, keepdim=True) + eps)       # inputs are zero-mean by construction (centred weights)


def template_stage1(sd, p, t, t_xyz, r, nhead=2):
    """per-object: M (B, 64 d, 64 out) head-split rows of blockdiag(KV) Wm_c^T (times 1/N), ksum (B, 64) (times 1/N)"""
    N = t.shape[1]
    pos = F.linear(F.relu(F.linear(t_xyz, sd[p + ".pos_mlp.0.weight"], sd[p + ".pos_mlp.0.bias"])), sd[p + ".pos_mlp.2.weight"], sd[p + ".pos_mlp.2.bias"])
    k = elu1(F.linear(t, sd[p + ".k_proj.weight"]))
    v = F.linear(t + pos, sd[p + ".v_proj.weight"])
    return kv_to_operand(k, v, center(sd[p + ".merge.weight"]), 1.0 / N, r, "MK1", None, nhead)


def kv_to_operand(Kf, V, Wm_c, scale, r, site, site_kvb, nhead):
    B, N, C = Kf.shape
    dh = C // nhead
    KV = torch.einsum("bnhd,bnhv->bhdv", Kf.view(B, N, nhead, dh), V.view(B, N, nhead, dh)) * scale      # (B, H, dh, dh)
    if site_kvb:
        KV = r(KV, site_kvb)
    ksum = Kf.sum(1) * scale                                                                            # (B, C)
    # M[b, h*dh + d, out] = sum_v KV[b,h,d,v] Wm[out, h*dh + v]
    M = torch.einsum("bhdv,ohv->bhdo", KV, Wm_c.view(C, nhead, dh)).reshape(B, C, C)
    return r(M, site), r(ksum, site)


def attend(Qf, M, ksum, eps, nhead=2):
    """Qf (P, N, C), M (P, C, C), ksum (P, C) -> merged message (P, N, C) before LayerNorm1"""
    P, N, C = Qf.shape
    dh = C // nhead
    out = 0
    for h in range(nhead):
        q = Qf[:, :, h * dh:(h + 1) * dh]
        num = torch.einsum("pnd,pdo->pno", q, M[:, h * dh:(h + 1) * dh])
        den = torch.einsum("pnd,pd->pn", q, ksum[:, h * dh:(h + 1) * dh]) + eps
        out = out + num / den.unsqueeze(-1)
    return out


def stage1(sd, s, M, ksum, r, eps):
    """s (P, N, C) search features, template operand (M, ksum) per pair -> a (P, N, C)"""
    p = "cross_stage1"
    d = s.shape[-1]
    Qf = r(elu1(F.linear(s, sd[p + ".q_proj.weight"])), "QF1")
    X = r(ln_noaffine(attend(Qf, M, ksum, eps)), "X1")
    W0 = sd[p + ".mlp.0.weight"]
    g1, b1, g2, b2 = sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], sd[p + ".norm2.weight"], sd[p + ".norm2.bias"]
    W0b = r(W0[:, d:] * g1[None, :], "W1")
    W0a = r(W0[:, :d], "W1")
    bias = r(W0[:, d:] @ b1 - W0[:, :d] @ b2, "W1")
    Hop = r(s + b2, "Hop")
    hid = r(F.relu(F.linear(X, W0b) + F.linear(Hop, W0a) + bias), "Hd1")
    y = F.linear(hid, r(center(sd[p + ".mlp.2.weight"]), "W1"))
    a32 = r(s + b2, "Hres") + g2 * ln_noaffine(y)
    return r(a32, "a"), r(a32, "a_res")            # operand image | what the stage-2 residual adds (a_res off: hi + lo images)


def stage2(sd, a, a_res, a_templ, pos_v, r, eps_raw, nhead=2):
    """a (P, N, C): search; a_templ (P, N, C): the other direction's stage-1 output; pos_v = Wv pos of the template (P, N, C)"""
    p = "cross_stage2"
    d = a.shape[-1]
    N = a_templ.shape[1]
    Kf = r(elu1(F.linear(a_templ, r(sd[p + ".k_proj.weight"], "W1b"))), "KfV")
    V = r(F.linear(a_templ, r(sd[p + ".v_proj.weight"], "W1b")) + r(pos_v, "PV"), "KfV")
    M, ksum = kv_to_operand(Kf, V, r(center(sd[p + ".merge.weight"]), "W1b"), 1.0 / N, r, "B7", "KVb", nhead)
    Qf = r(elu1(F.linear(a, r(sd[p + ".q_proj.weight"], "W2"))), "Qf2")
    X = r(ln_noaffine(attend(Qf, M, ksum, eps_raw / N)), "X2")
    W0 = sd[p + ".mlp.0.weight"]
    g1, b1, g2, b2 = sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], sd[p + ".norm2.weight"], sd[p + ".norm2.bias"]
    hid = r(F.relu(F.linear(a, r(W0[:, :d], "W2")) + F.linear(X, r(W0[:, d:] * g1[None, :], "W2")) + r(W0[:, d:] @ b1, "W2")), "Hd2")
    y = F.linear(hid, r(center(sd[p + ".mlp.2.weight"]), "W2"))
    return a_res + g2 * ln_noaffine(y) + b2


@torch.no_grad()
def logits(orc, O, h_t, xyz_t, h_d, xyz_d, r, chunk=1024):
    sd = orc.sd
    T, D = h_t.shape[0], h_d.shape[0]
    N = h_t.shape[2]
    ht, hd = h_t.permute(0, 2, 1).contiguous(), h_d.permute(0, 2, 1).contiguous()
    Mt, kst = template_stage1(sd, "cross_stage1", ht, xyz_t, r)
    Md, ksd = template_stage1(sd, "cross_stage1", hd, xyz_d, r)
    p2 = "cross_stage2"
    posv = lambda xyz: F.linear(F.linear(F.relu(F.linear(xyz, sd[p2 + ".pos_mlp.0.weight"], sd[p2 + ".pos_mlp.0.bias"])), sd[p2 + ".pos_mlp.2.weight"],
                                         sd[p2 + ".pos_mlp.2.bias"]), sd[p2 + ".v_proj.weight"])
    pvt, pvd = posv(xyz_t), posv(xyz_d)
    pairs = torch.cartesian_prod(torch.arange(T), torch.arange(D))
    out = torch.zeros(T, D)
    eps = 1e-6
    for s0 in range(0, pairs.shape[0], chunk):
        pr = pairs[s0:s0 + chunk]
        i, j = pr[:, 0], pr[:, 1]
        a, a_res = stage1(sd, ht[i], Md[j], ksd[j], r, eps / N)
        b, b_res = stage1(sd, hd[j], Mt[i], kst[i], r, eps / N)
        o1 = stage2(sd, a, a_res, b, pvd[j], r, eps)
        o2 = stage2(sd, b, b_res, a, pvt[i], r, eps)
        o = torch.cat([o1, o2], dim=1).permute(0, 2, 1)           # point-cat: (P, C, 2N)
        out[i, j] = orc._head(O.pooled_feats(o, "both"))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fmt", default="f16")
    ap.add_argument("--tracks", type=int, default=48)
    ap.add_argument("--dets", type=int, default=128)
    ap.add_argument("--npts", type=int, default=256)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    import helpers
    from oracle import reid_oracle as O
    torch.set_num_threads(os.cpu_count())
    _, orc = helpers.build_pair("pt", (256, 128, 64), device="cpu", perturb=False)
    t, d = O.synth_objects(a.tracks, a.npts, 1000), O.synth_objects(a.dets, a.npts, 1)
    xt, ht = orc.encode(t)
    xd, hd = orc.encode(d)
    ref = orc.match_all_pairs(ht, xt, hd, xd, chunk=2048)
    base = logits(orc, O, ht, xt, hd, xd, Rounder(a.fmt, []))
    print(f"emulation vs oracle in fp32: max {float((base - ref).abs().max()):.2e}  (logit std {float(ref.std()):.4f})")

    def report(name, on):
        L = logits(orc, O, ht, xt, hd, xd, Rounder(a.fmt, on))
        e = L - ref
        ec = e - e.mean(1, keepdim=True)             # what can change a row's arg-max: the error with the row's common shift removed
        top2 = torch.topk(ref, 2, dim=1)[0]
        gap = (top2[:, 0] - top2[:, 1])
        # expected flips: a row flips when the error difference between its two best columns exceeds their gap
        i1 = ref.argmax(1)
        print(f"{name:28s} rms {float(e.pow(2).mean().sqrt()):.2e}  row-centred rms {float(ec.pow(2).mean().sqrt()):.2e}  max {float(e.abs().max()):.2e}  "
              f"top1 {float((L.argmax(1) == i1).float().mean()):.4f}", flush=True)

    report("all sites", SITES)
    if a.only:
        for grp in a.only.split(";"):
            report("only " + grp, grp.split(","))
            report("all but " + grp, [s for s in SITES if s not in grp.split(",")])
        return
    for s in SITES:
        report("only " + s, [s])
    for grp in (["Hres", "a"], ["W1", "W1b", "W2"], ["Hres", "a", "Hop"]):
        report("all but " + ",".join(grp), [s for s in SITES if s not in grp])


if __name__ == "__main__":
    main()
