#!/usr/bin/env python
"""Derive the per-unit DRAM traffic of the fused pair kernels from an `ncu --set full` capture and write the file bench.py
reads for `roofline.traffic` (profiles/rNN_ncu_traffic.json) -- so that the number in the bench line always comes from a
capture of the kernels, never from a literal in bench.py.

  python scripts/ncu_traffic.py --out profiles/r02_ncu_traffic.json --mode parity_tc --rep gpurun_out/x.ncu-rep --units 65536 \
                                [--mode fast --rep ... --units ...]
units = (pair, direction) units per launch of the captured command (tracks x dets of the capture run when it is one chunk).
"""
import argparse
import csv
import json
import os
import re
import subprocess
import sys


def kernels_of(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(raw.splitlines()))
    hdr = rd[0]
    col = {h: i for i, h in enumerate(hdr)}
    out = {}
    for r in rd[2:]:
        if len(r) < len(hdr):
            continue
        m = re.search(r"(\w+)\s*(?:<[^()]*>)?\s*\(", r[col["Kernel Name"]].replace("<unnamed>::", ""))
        name = m.group(1) if m else r[col["Kernel Name"]]
        g = lambda m: float(r[col[m]].replace(",", "")) if m in col and r[col[m]] else 0.0
        unit = lambda m: rd[1][col[m]] if m in col else ""
        rdb, wrb = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rdb *= scale.get(unit("dram__bytes_read.sum"), 1.0)
        wrb *= scale.get(unit("dram__bytes_write.sum"), 1.0)
        t = g("gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(unit("gpu__time_duration.sum"), 1e-6)
        e = out.setdefault(name, {"launches": 0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "ms": 0.0, "tensor_pipe_pct": [],
                                  "registers": g("launch__registers_per_thread")})
        e["launches"] += 1
        e["dram_read_bytes"] += rdb
        e["dram_write_bytes"] += wrb
        e["ms"] += t
        e["tensor_pipe_pct"].append(g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--mode", action="append", required=True)
    ap.add_argument("--rep", action="append", required=True)
    ap.add_argument("--units", action="append", type=int, required=True)
    ap.add_argument("--command", default="")
    a = ap.parse_args()
    res = {"how": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per launch / units per launch "
                  "(cold-cache, serialised replays); written by scripts/ncu_traffic.py", "command": a.command, "modes": {}}
    for mode, rep, units in zip(a.mode, a.rep, a.units):
        ks = kernels_of(rep)
        ent = {"report": os.path.basename(rep), "units_per_launch": units, "kernels": {}}
        for n, e in ks.items():
            L = e["launches"]
            ent["kernels"][n] = {"launches_captured": L, "dram_bytes_per_launch": (e["dram_read_bytes"] + e["dram_write_bytes"]) / L,
                                 "dram_read_bytes_per_launch": e["dram_read_bytes"] / L, "dram_write_bytes_per_launch": e["dram_write_bytes"] / L,
                                 "dram_bytes_per_unit": (e["dram_read_bytes"] + e["dram_write_bytes"]) / L / units,
                                 "ms_per_launch_under_ncu": e["ms"] / L, "tensor_pipe_active_pct": sum(e["tensor_pipe_pct"]) / L,
                                 "registers_per_thread": e["registers"]}
        ent["dram_bytes_per_pair"] = 2 * sum(k["dram_bytes_per_unit"] for k in ent["kernels"].values())
        res["modes"][mode] = ent
    json.dump(res, open(a.out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    sys.exit(main())
