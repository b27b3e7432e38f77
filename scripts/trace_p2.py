"""Cycle trace of one group of pair_p2_kernel (debug): prints the average cycles between stage markers."""
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
_lib.lib().pcreid_pair_tc_set_trace(ctypes.c_void_p(buf.data_ptr()))
m.match_all_pairs(ht, xt, hd, xd)
torch.cuda.synchronize()
_lib.lib().pcreid_pair_tc_set_trace(None)
b = buf.cpu().tolist()
ev = [(b[i], b[i + 1]) for i in range(0, 2040, 2) if b[i + 1] != 0]
names = {100: "tile start", 101: "after load-wait+sync", 102: "after R1 copy+publish+prefetch issue", 103: "G4' issued", 104: "G4' done(wait)",
         105: "Qf epilogue", 106: "publish", 107: "G7 issued", 108: "G7 done", 109: "attn+LN epilogue", 110: "publish", 111: "G8 issued",
         112: "G8 done", 113: "relu128 epilogue", 114: "publish", 115: "G9 issued", 116: "G9 done", 200: "elected", 201: "8 MMAs issued", 202: "commit issued", 117: "LN+res+transpose store", 118: "sync"}
agg = collections.defaultdict(list)
for (c0, t0), (c1, t1) in zip(ev[:-1], ev[1:]):
    agg[(t0, t1)].append(c1 - c0)
tot = 0
for (t0, t1), v in sorted(agg.items()):
    if len(v) < 3: continue
    v = sorted(v); med = v[len(v) // 2]
    print(f"{t0}->{t1} {names.get(t1, ''):40s} median {med:6d} cyc  (n={len(v)})")
    tot += med
print("sum of medians per tile", tot)
