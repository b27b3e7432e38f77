#!/usr/bin/env python
"""Summarise the SASS source page of one kernel of an ncu report: stall reasons, opcode histogram (warp instructions and stall
samples), hottest instructions.  python scripts/ncu_source_summary.py rep.ncu-rep regex:kernel [launch-skip]"""
import collections
import csv
import re
import subprocess
import sys


def main(rep, kname, skip="0", top=14):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kname, "--launch-skip", skip, "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print(rows[hi - 1][1] if hi else "")
    hdr = rows[hi]
    c = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
    iv = lambda r, h: int(float(r[c[h]] or 0)) if h in c else 0
    tot = sum(iv(r, "# Samples") for r in data)
    inst = sum(iv(r, "Instructions Executed") for r in data)
    print(f"samples {tot}, warp instructions {inst}, SASS lines {len(data)}")
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(iv(r, h) for r in data) for h in reasons}
    print("stalls:", ", ".join(f"{k[6:]} {v / max(tot, 1):.3f}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    op, ops = collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[c["Source"]])
        if m:
            o = m.group(2).split(".")[0]
            op[o] += iv(r, "Instructions Executed")
            ops[o] += iv(r, "# Samples")
    print("opcode: share of warp instructions / share of stall samples")
    for k, v in op.most_common(22):
        print(f"  {k:10s} {v / max(inst, 1):.3f} / {ops[k] / max(tot, 1):.3f}")
    print("hottest instructions:")
    for r in sorted(data, key=lambda r: -iv(r, "# Samples"))[:top]:
        why = max(reasons, key=lambda h: iv(r, h))
        print(f"  {iv(r, '# Samples') / max(tot, 1):.3f} {why[6:]:12s} {r[c['Source']].strip()[:110]}")


if __name__ == "__main__":
    main(*sys.argv[1:4])
