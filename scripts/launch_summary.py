#!/usr/bin/env python
"""Condense an ncu launch list (`--metrics gpu__time_duration.sum --csv`) into a per-kernel table: launches, total / mean
duration, share of the captured GPU time.  python scripts/launch_summary.py gpurun_out/x.csv > profiles/rNN_launches_x.md"""
import csv
import re
import sys


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    c = {h: i for i, h in enumerate(hdr)}
    agg = {}
    for r in rows:
        if r[c["Metric Name"]] != "gpu__time_duration.sum":
            continue
        full = r[c["Kernel Name"]].replace("<unnamed>::", "")
        m = re.search(r"([\w:]+)\s*(<[^()]*>)?\s*\(", full)
        name = (m.group(1).split("::")[-1] + (m.group(2) or "")) if m else full
        name = name[:70]
        ns = float(r[c["Metric Value"]].replace(",", ""))
        ns *= {"ns": 1, "us": 1e3, "ms": 1e6}.get(r[c["Metric Unit"]], 1)
        e = agg.setdefault(name, [0, 0.0, r[c["Block Size"]], r[c["Grid Size"]]])
        e[0] += 1
        e[1] += ns
    tot = sum(e[1] for e in agg.values())
    print(f"source: {path}; {sum(e[0] for e in agg.values())} launches, {tot / 1e6:.3f} ms of GPU time (per-launch times under ncu are cold-cache and serialised: read the SHARES)\n")
    print("| kernel | launches | total ms | mean us | share | block | grid (last) |\n|---|---|---|---|---|---|---|")
    for n, e in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {e[0]} | {e[1] / 1e6:.3f} | {e[1] / e[0] / 1e3:.1f} | {e[1] / tot:.3f} | {e[2]} | {e[3]} |")


if __name__ == "__main__":
    main(sys.argv[1])
