#!/usr/bin/env python
"""Which encoder kernels cost the tensor-core parity mode its top-1 agreement?  Runs the PT encoder with the tcgen05 (tf32)
kernels enabled per module class -- SA shared MLPs, Self_Attention blocks, FP_SA blocks -- scores with the fp32 matcher and
with the fp16 matcher, and reports embedding error, logit error and RAW top-1 agreement against the CPU oracle.

  python scripts/encoder_error_probe.py [--rows 256] [--cols 256]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=256)
    ap.add_argument("--cols", type=int, default=256)
    ap.add_argument("--seeds", default="1000:1,0:1")
    args = ap.parse_args()
    import helpers
    from oracle import reid_oracle as O
    from pcreid_b200.models import pointnet2_utils as PU
    N = 256
    m, orc = helpers.build_pair("pt", (N, N // 2, N // 4), device="cuda", perturb=False)
    torch.set_num_threads(os.cpu_count())
    classes = {"sa": PU.PointNetSetAbstractionEdgeSA, "self": PU.Self_Attention, "fp": PU.FP_SA}
    for seeds in args.seeds.split(","):
        st, sd_ = (int(x) for x in seeds.split(":"))
        t, d = O.synth_objects(args.rows, N, st), O.synth_objects(args.cols, N, sd_)
        oxt, oht = orc.encode(t)
        oxd, ohd = orc.encode(d)
        Lo = orc.match_all_pairs(oht, oxt, ohd, oxd, chunk=4096)
        for combo in ("", "sa", "self", "fp", "sa,self", "sa,fp", "self,fp", "sa,self,fp"):
            on = set(combo.split(",")) if combo else set()
            m.set_mode("parity_tc")
            for mod in m.modules():
                for name, cls in classes.items():
                    if type(mod) is cls and hasattr(mod, "tc_mode"):
                        mod.tc_mode = name in on
            xt, ht = m.encode(t.cuda())
            xd, hd = m.encode(d.cuda())
            out = {"seeds": seeds, "tc": combo or "none", "enc_rms": float((ht.cpu() - oht).pow(2).mean().sqrt())}
            for mat in ("parity", "parity_tc"):
                m.match_mode = mat
                L = m.match_all_pairs(ht, xt, hd, xd).cpu()
                e = L - Lo
                out[mat] = {"rms": float(e.pow(2).mean().sqrt()), "max": float(e.abs().max()), "top1": float((L.argmax(1) == Lo.argmax(1)).float().mean())}
            print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
