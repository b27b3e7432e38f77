"""The five BASELINE.json configs (C1..C5) on one B200: timings of encode / match per mode, and parity of a
sampled sub-block against the oracle.  Bench lines are produced by bench.py (C2); this script is the evidence that the
other configurations run at their named shapes.  Writes JSON to stdout."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, helpers
from oracle import reid_oracle as O
dev = "cuda"


def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e


def run(name, kind, N, blist, T, D, mode, check=8, reps=2):
    m, orc = helpers.build_pair(kind, blist, device=dev, perturb=False)
    m.set_mode(mode)
    t, d = O.synth_objects(T, N, 0), O.synth_objects(D, N, 1)
    td, dd = t.to(dev), d.to(dev)
    for _ in range(reps):
        e0 = ev(); xt, ht = m.encode(td); xd, hd = m.encode(dd); e1 = ev()
        L = m.match_all_pairs(ht, xt, hd, xd); e2 = ev(); torch.cuda.synchronize()
    enc, mat = e0.elapsed_time(e1), e1.elapsed_time(e2)
    oxt, oht = orc.encode(t[:check]); oxd, ohd = orc.encode(d[:check])
    Lo = orc.match_all_pairs(oht, oxt, ohd, oxd)
    err = float((L[:check, :check].cpu() - Lo).abs().max())
    r = {"config": name, "backbone": kind, "points": N, "tracks": T, "dets": D, "mode": mode, "encode_ms": enc, "match_ms": mat,
         "objects_per_s": (T + D) / enc * 1e3, "pairs_per_s": T * D / mat * 1e3, "max_abs_dlogit_vs_oracle_on_sample": err,
         "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}
    print(json.dumps(r), flush=True)
    torch.cuda.reset_peak_memory_stats()
    return r


def run_image(name, S, T, D, mode, check=6):
    """token side of ImageReIDNet (reid_image_deit-tiny_point-cat.py): downsample of (T + D) x S tokens x 192 + all-pairs match"""
    m, orc = helpers.build_image_pair(device=dev, perturb=False)
    m.set_mode(mode)
    raw = O.synth_tokens(T + D, 192, S, 0)
    rd = raw.to(dev)
    for _ in range(2):
        e0 = ev(); h = m.downsample_tokens(rd); e1 = ev()
        L = m.match_all_pairs(h[:T], h[T:]); e2 = ev(); torch.cuda.synchronize()
    enc, mat = e0.elapsed_time(e1), e1.elapsed_time(e2)
    ho = orc.downsample_tokens(raw)
    Lo = orc.match_all_pairs(ho[:check], ho[T:T + check])
    err = float((L[:check, :check].cpu() - Lo).abs().max())
    print(json.dumps({"config": name, "backbone": "image tokens (DeiT-tiny shape)", "tokens": S, "tracks": T, "dets": D, "mode": mode,
                      "downsample_ms": enc, "match_ms": mat, "pairs_per_s": T * D / mat * 1e3,
                      "max_abs_dlogit_vs_oracle_on_sample": err, "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}), flush=True)
    torch.cuda.reset_peak_memory_stats()


only = os.environ.get("ONLY", "")
jobs = [
    ("c1", lambda: run("C1 PointNet 64x64 @128", "pointnet", 128, (128, 64, 32), 64, 64, "parity")),
    ("c1", lambda: run("C1 PointNet 64x64 @128", "pointnet", 128, (128, 64, 32), 64, 64, "fast")),
    ("c2", lambda: run("C2 PT 1024x1024 @256", "pt", 256, (256, 128, 64), 1024, 1024, "fast")),
    ("c3", lambda: run("C3 DGCNN 2048x2048 @256", "dgcnn", 256, (256, 128, 64), 2048, 2048, "fast", reps=1)),
    ("c4", lambda: run("C4 PT 512x4096 @1024 (one rank's row block of 4096x4096 over 8 GPUs)", "pt", 1024, (1024, 512, 256), 512, 4096,
                       "fast", check=4, reps=1)),
    ("c5", lambda: run("C5 sweep PT 256x256 @256", "pt", 256, (256, 128, 64), 256, 256, "fast")),
    ("c5", lambda: run("C5 sweep PT 1024x1024 @256", "pt", 256, (256, 128, 64), 1024, 1024, "fast")),
    ("c5", lambda: run("C5 sweep PT 4096x4096 @256", "pt", 256, (256, 128, 64), 4096, 4096, "fast", reps=1)),
    # the 'concat' baseline config pools with MaxPool1d(64) over the channel axis -> width N per object: only consistent at N=128
    ("concat", lambda: run("C5 concat head PT 4096x4096 @128", "concat", 128, (128, 64, 32), 4096, 4096, "parity")),
    ("concat", lambda: run("C5 concat head PT 4096x4096 @128 (tensor-core head)", "concat", 128, (128, 64, 32), 4096, 4096, "fast")),
    ("concat", lambda: run("target shape: concat head PT 16384x16384 @128 (tensor-core head)", "concat", 128, (128, 64, 32), 16384, 16384, "fast", reps=1)),
    ("concat", lambda: run("target shape: concat head PT 4096x4096 @128, 16384x16384", "concat", 128, (128, 64, 32), 16384, 16384, "parity", reps=1)),
]
jobs += [
    ("image", lambda: run_image("image-token matcher 256x256 @198 tokens", 198, 256, 256, "parity")),
    ("image", lambda: run_image("image-token matcher 1024x1024 @198 tokens", 198, 1024, 1024, "fast")),
]
# model variants the fused xcorr_eff matcher does not cover (d_model = 128 / widened SA: -7M / -1.5M configs; 'xcorr' with the local
# attention stages; 'xcorr-baseline'): unfused chains, fp32 FFMA kernels in 'parity', TMA-staged tf32 tcgen05 GEMMs in 'fast'
for kind in ("pt7m", "pt15m", "xcorr", "xcorr-baseline"):
    for mode in ("parity", "fast"):
        jobs.append(("variants", (lambda k=kind, mo=mode: run(f"variant {k} 256x256 @256", k, 256, (256, 128, 64), 256, 256, mo, check=6, reps=2))))
for tag, job in jobs:
    if not only or tag in only.split(","):
        job()
