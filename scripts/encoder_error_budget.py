#!/usr/bin/env python
"""CPU study: which Linear layers of the PT encoder's attention blocks cost the tf32 mode its top-1 agreement?
The oracle's `_lin` is wrapped so that, for layer names matching a pattern, input and weight are rounded to tf32 (10-bit
mantissa, round-to-nearest) before the fp32 product -- the arithmetic of a kind::tf32 MMA with pre-rounded operands; the
linear-attention einsums are rounded with the block they belong to.  Logits come from the fp32 oracle matcher, so only the
encoder differs.  Reports embedding rms error, logit rms / row-centred rms error and raw top-1 agreement per pattern.

  python scripts/encoder_error_budget.py [--rows 128] [--cols 128]
"""
import argparse
import os
import re
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def tf32(x):
    b = x.contiguous().view(torch.int32)
    return ((b + 0x1000) & ~0x1fff).view(torch.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=128)
    ap.add_argument("--cols", type=int, default=128)
    ap.add_argument("--patterns", default="")
    a = ap.parse_args()
    import helpers
    from oracle import reid_oracle as O
    torch.set_num_threads(os.cpu_count())
    _, orc = helpers.build_pair("pt", (256, 128, 64), device="cpu", perturb=False)
    t, d = O.synth_objects(a.rows, 256, 1000), O.synth_objects(a.cols, 256, 1)
    xt, ht = orc.encode(t)
    xd, hd = orc.encode(d)
    ref = orc.match_all_pairs(ht, xt, hd, xd, chunk=2048)
    lin0 = O._lin
    state = {"pat": None, "act_only": False}

    def lin(sd, p, x):
        if state["pat"] is not None and re.search(state["pat"], p):
            w = sd[p + ".weight"] if state["act_only"] else tf32(sd[p + ".weight"])
            return F.linear(tf32(x), w, sd.get(p + ".bias"))
        return lin0(sd, p, x)

    O._lin = lin
    pats = a.patterns.split(";") if a.patterns else [
        r"FP_modules\.", r"FP_modules\.0", r"FP_modules\.1", r"FP_modules\.2", r"SA_modules\.\d\.(?!mlp)", r"FP_modules\.0.*q_proj", r"FP_modules\.0.*mlp\.0",
        r"FP_modules\.0.*mlp\.2", r"FP_modules\.0.*merge", r"FP_modules\.0.*(k_proj|v_proj|pos)", r"FP_modules\.0.*(q_proj|mlp\.0)"]
    for act_only in (False, True):
        for pat in pats:
            state["pat"], state["act_only"] = pat, act_only
            x1, h1 = orc.encode(t)
            x2, h2 = orc.encode(d)
            L = orc.match_all_pairs(h1, x1, h2, x2, chunk=2048)
            e = L - ref
            ec = e - e.mean(1, keepdim=True)
            print(f"{'act-only ' if act_only else 'act+weight '}{pat:42s} enc rms {float((h1 - ht).pow(2).mean().sqrt()):.2e}  logit rms {float(e.pow(2).mean().sqrt()):.2e}  "
                  f"row-centred {float(ec.pow(2).mean().sqrt()):.2e}  top1 {float((L.argmax(1) == ref.argmax(1)).float().mean()):.4f}", flush=True)


if __name__ == "__main__":
    main()
