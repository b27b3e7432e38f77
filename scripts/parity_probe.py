#!/usr/bin/env python
"""Error budget of the tensor-core modes against the CPU oracle on a block of the C2 workload (BASELINE configs[1]:
PT, 256 pts): for every (encoder mode, matcher mode) combination report max / mean |dlogit|, the RAW row top-1 agreement
and the number of decisive rows (oracle gap > 2 x max error).  Test infrastructure (imports oracle/).

  python scripts/parity_probe.py [--rows 128] [--cols 256] [--modes parity,fast,...]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402


def stats(Lo, L):
    err = (L - Lo).abs()
    top2 = torch.topk(Lo, 2, dim=1)[0]
    gap = top2[:, 0] - top2[:, 1]
    same = Lo.argmax(1) == L.argmax(1)
    mx = float(err.max())
    return {"max_abs": mx, "mean_abs": float(err.mean()), "rms": float(err.pow(2).mean().sqrt()),
            "top1_raw": float(same.float().mean()), "decisive_rows": int((gap > 2 * mx).sum()),
            "decisive_ok": bool(same[gap > 2 * mx].all()), "rows": int(Lo.shape[0]), "cols": int(Lo.shape[1]),
            "flipped_gap_max": float(gap[~same].max()) if (~same).any() else 0.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=128)
    ap.add_argument("--cols", type=int, default=256)
    ap.add_argument("--npts", type=int, default=256)
    ap.add_argument("--combos", default="parity:parity,fast:parity,parity:fast,parity:parity_tc,fast:fast,parity_tc:parity_tc")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import helpers
    from oracle import reid_oracle as O
    N = args.npts
    bl = (N, N // 2, N // 4)
    m, orc = helpers.build_pair("pt", bl, device="cuda", perturb=False)
    t, d = O.synth_objects(args.rows, N, 0), O.synth_objects(args.cols, N, 1)
    torch.set_num_threads(os.cpu_count())
    t0 = time.perf_counter()
    oxt, oht = orc.encode(t)
    oxd, ohd = orc.encode(d)
    t1 = time.perf_counter()
    Lo = orc.match_all_pairs(oht, oxt, ohd, oxd, chunk=4096)
    t2 = time.perf_counter()
    top2 = torch.topk(Lo, 2, dim=1)[0]
    gap = top2[:, 0] - top2[:, 1]
    res = {"oracle": {"encode_s": t1 - t0, "match_s": t2 - t1, "logit_std": float(Lo.std()), "gap_median": float(gap.median()),
                      "gap_min": float(gap.min()), "gap_p10": float(gap.kthvalue(max(1, args.rows // 10))[0])}}
    print(json.dumps(res["oracle"]), flush=True)
    for combo in args.combos.split(","):
        enc, mat = combo.split(":")
        m.set_mode(enc)
        xt, ht = m.encode(t.cuda())
        xd, hd = m.encode(d.cuda())
        m.set_mode(mat)
        L = m.match_all_pairs(ht, xt, hd, xd).cpu()
        s = stats(Lo, L)
        s["enc_err_max"] = float(max((ht.cpu() - oht).abs().max(), (hd.cpu() - ohd).abs().max()))
        s["enc_err_rms"] = float((ht.cpu() - oht).pow(2).mean().sqrt())
        res[combo] = s
        print(combo, json.dumps(s), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
