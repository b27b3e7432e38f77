"""Cycle trace of one group (group 0 of CTA 0) of pair_p2y_kernel (debug): median cycles between stage markers."""
import collections, ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, helpers
from oracle import reid_oracle as O
from pcreid_b200 import _lib
dev = "cuda"
m, _ = helpers.build_pair("pt", (256, 128, 64), device=dev)
m.set_mode('fast')
T = D = 256
t, d = O.synth_objects(T, 256, 0).to(dev), O.synth_objects(D, 256, 1).to(dev)
xt, ht = m.encode(t); xd, hd = m.encode(d)
m.match_all_pairs(ht, xt, hd, xd)
buf = torch.zeros(2048, dtype=torch.int64, device=dev)
_lib.lib().pcreid_pair_tc2_set_trace(ctypes.c_void_p(buf.data_ptr()))
m.match_all_pairs(ht, xt, hd, xd)
torch.cuda.synchronize()
_lib.lib().pcreid_pair_tc2_set_trace(None)
b = buf.cpu().tolist()
ev = [(b[i], b[i + 1]) for i in range(0, 2040, 2) if b[i] != 0]
names = {0: "prev tile end -> tile start", 1: "cp.async wait + publish", 2: "Gq issue..done", 3: "Qf epilogue", 4: "publish", 5: "G7 issue..done",
         6: "attention epilogue", 7: "publish", 8: "G8 issue + side loads .. done", 9: "Hd epilogue (ld, pack.relu, st)", 10: "fence + sync",
         11: "G9 issue..done", 12: "LN2 + residual + pooling"}
agg = collections.defaultdict(list)
for (c0, t0), (c1, t1) in zip(ev[:-1], ev[1:]):
    agg[(t0, t1)].append(c1 - c0)
tot = 0
for (t0, t1), v in sorted(agg.items()):
    if len(v) < 3: continue
    v = sorted(v); med = v[len(v) // 2]
    print(f"{t0:3d}->{t1:3d} {names.get(t1, ''):42s} median {med:6d} cyc  p10 {v[len(v) // 10]:6d}  p90 {v[len(v) * 9 // 10]:6d} (n={len(v)})")
    tot += med
print("sum of medians per tile", tot)
