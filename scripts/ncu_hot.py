"""Summarise an ncu report: per-kernel key metrics, stall reasons and the hottest SASS lines."""
import collections, csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rd = list(csv.reader(raw.splitlines()))
hdr = rd[0]
want = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__cycles_active.avg', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__registers_per_thread',
        'smsp__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_lsu.sum', 'smsp__cycles_active.avg']
r = rd[2]
for w in want:
    if w in hdr: print(f"{w:75s} {r[hdr.index(w)]} {rd[1][hdr.index(w)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rd = list(csv.reader(src.splitlines()))
hdr = rd[1]
iS, iN, iX = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
rows, seen = [], set()
for x in rd[2:]:
    if len(x) <= max(stall) or not x[iN].isdigit(): continue
    if x[0] in seen: break
    seen.add(x[0]); rows.append(x)
tot = sum(int(x[iN]) for x in rows)
agg = collections.Counter()
for x in rows:
    for i in stall: agg[hdr[i]] += int(x[i] or 0)
print("samples", tot, "| instr executed (warp)", sum(int(x[iX]) for x in rows))
print("stalls:", ", ".join(f"{k[6:]} {v/tot:.1%}" for k, v in agg.most_common(9)))
for x in sorted(rows, key=lambda x: -int(x[iN]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 22]:
    top = max(stall, key=lambda i: int(x[i] or 0))
    print(x[iN].rjust(6), x[iX].rjust(9), hdr[top][6:].ljust(12), x[iS].strip()[:96])
