#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json: "pair scores/sec + objects encoded/sec
(PT, 256 pts)"), configs[1]: Point Transformer (pts_point-transformer_r_nus_det) inference, 1024 tracks +
1024 detections x 256 points, 1024 x 1024 all-pairs `xcorr_eff` match per GPU.

A step = encode the step's tracks and detections (PT backbone) + score every track x detection pair.
At N GPUs the track rows are sharded (weak scaling: 1024 tracks per rank, the 1024 detections are
split for encoding and all-gathered once).  Prints ONE JSON line (rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode parity]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

T_PER_GPU, D_TOTAL, NPTS = 1024, 1024, 256
BLIST = (256, 128, 64)
# algorithmic work of the reference formulation (torch.utils.flop_counter on the reference)
FLOP_PER_OBJECT = 592e6          # Pointnet_Backbone @256 pts
FLOP_PER_PAIR = 101.25e6         # xcorr_eff + pool + head @256 pts
METRIC = "pair scores/sec (PT-256 encode + all-pairs xcorr_eff match)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 8 and r[1].replace('.', '').isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 8 and r[2].replace('.', '').isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_reference_sample(n_obj=32, n_pairs=2048, threads=None):
    """The reference's PyTorch path (oracle restatement, bit-exact vs the reference modules) on host cores, on a
    bounded sample of the same workload: encode n_obj+n_obj objects, score n_pairs pairs; extrapolated to the
    step's composition (objects and pairs per step)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers                                  # oracle builder shared with the parity tests
    from oracle import reid_oracle as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    _, orc = helpers.build_pair("pt", BLIST, device="cpu", perturb=False)
    t, d = O.synth_objects(n_obj, NPTS, 0), O.synth_objects(n_obj, NPTS, 1)
    t0 = time.perf_counter()
    xt, ht = orc.encode(t)
    xd, hd = orc.encode(d)
    t1 = time.perf_counter()
    rows = max(1, n_pairs // n_obj)
    orc.match_all_pairs(ht[:rows], xt[:rows], hd, xd, chunk=4096)
    t2 = time.perf_counter()
    s_obj = (t1 - t0) / (2 * n_obj)
    s_pair = (t2 - t1) / (rows * n_obj)
    step_s = s_obj * (T_PER_GPU + D_TOTAL) + s_pair * T_PER_GPU * D_TOTAL
    return {"value": T_PER_GPU * D_TOTAL / step_s, "unit": "pairs/s", "cores": threads, "kind": "port",
            "sample": f"{2 * n_obj} objects encoded ({1 / s_obj:.1f} obj/s) + {rows * n_obj} pairs scored ({1 / s_pair:.1f} pairs/s), "
                      f"extrapolated to {T_PER_GPU + D_TOTAL} objects + {T_PER_GPU * D_TOTAL} pairs per step",
            "objects_per_s": 1 / s_obj, "pairs_only_per_s": 1 / s_pair}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    last = None
    for i in range(args.warmup + args.steps):
        last = cpu_reference_sample(32, 4096)
        if i >= args.warmup:
            vals.append(last["value"])
    v = sum(vals) / len(vals)
    last["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * T_PER_GPU * D_TOTAL / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"PT encode {T_PER_GPU}+{D_TOTAL} objects x {NPTS} pts + {T_PER_GPU}x{D_TOTAL} xcorr_eff pairs "
                                   "(reference PyTorch path on host cores, bounded sample per step, extrapolated)"},
            "cpu_baseline": last, "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="fast", choices=["parity", "fast"])
    ap.add_argument("--tracks", type=int, default=T_PER_GPU)
    ap.add_argument("--dets", type=int, default=D_TOTAL)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from pcreid_b200 import _lib, synthetic as S   # the measured legs never import oracle/ or tests/ (cpu_baseline leg only)
    from pcreid_b200.parallel import encode_and_gather, match_all_pairs_sharded, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    T_loc, D = args.tracks, args.dets
    d0, d1 = shard_range(D, rank, world)
    det_counts = [shard_range(D, r, world)[1] - shard_range(D, r, world)[0] for r in range(world)]

    torch.manual_seed(66)
    from pcreid_b200.models import build_model
    model = build_model(S.point_transformer_cfg(BLIST)).eval().to(dev)
    model.set_mode(args.mode)
    tracks_h = S.synth_objects(T_loc, NPTS, 1000 + rank).pin_memory()
    dets_h = S.synth_objects(D, NPTS, 1)[d0:d1].contiguous().pin_memory()
    tracks_d, dets_d = tracks_h.to(dev), dets_h.to(dev)
    out_h = torch.empty((T_loc, D), dtype=torch.float32).pin_memory()

    def step_device():
        return match_all_pairs_sharded(model, tracks_d, dets_d, det_counts)

    def step_e2e():
        t = tracks_h.to(dev, non_blocking=True)
        d = dets_h.to(dev, non_blocking=True)
        rows = match_all_pairs_sharded(model, t, d, det_counts)
        out_h.copy_(rows, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        barrier()
        ev[0].record()
        for i in range(steps):
            fn()
            ev[i + 1].record()
        barrier()
        ms = ev[0].elapsed_time(ev[steps])
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    calls0 = _lib.ABI_CALLS
    total_ms, per_step = timed(step_device, args.steps)
    launches = _lib.ABI_CALLS - calls0
    clocks = sampler.stop() if rank == 0 else None
    # phase split (device events, one extra step): encode vs match
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    torch.cuda.synchronize()
    e0.record()
    xt, ht, xd, hd = encode_and_gather(model, tracks_d, dets_d, det_counts)
    e1.record()
    model.match_all_pairs(ht, xt, hd, xd)
    e2.record()
    torch.cuda.synchronize()
    enc_ms, match_ms = e0.elapsed_time(e1), e0.elapsed_time(e2) - e0.elapsed_time(e1)
    # per-launch durations of the fused kernels (CUDA events on the launching stream), one more match pass
    kern = {}
    if args.mode == "fast" and model._fused is not None:
        model._fused.timing = []
        model.match_all_pairs(ht, xt, hd, xd)
        torch.cuda.synchronize()
        for name, a0, a1, units in model._fused.timing:
            k = kern.setdefault(name, {"launches": 0, "ms": 0.0, "units": 0})
            k["launches"] += 1
            k["ms"] += a0.elapsed_time(a1)
            k["units"] += units
        model._fused.timing = None
    for _ in range(1):
        step_e2e()
    e2e_ms, _ = timed(step_e2e, args.steps)

    if rank == 0:
        pk, pk_src = peaks()
        pairs_total = T_loc * world * D
        ms_step = total_ms / args.steps
        value = pairs_total / (ms_step * 1e-3)
        n_enc = T_loc + (d1 - d0)
        match_tflops = FLOP_PER_PAIR * T_loc * D / (match_ms * 1e-3) / 1e12 if match_ms > 0 else None
        peak_tf = pk["bf16_tflops_sustained"]
        roof_kernel = "match stage (cn_linear_kernel<*> dominates; unfused fp32 parity path)"
        roof_extra, roof_traffic = {}, None
        if kern:
            # dominant kernels: the two fused phases; algorithmic FLOPs of the reference formulation (101.25 MFLOP/pair:
            # stage 1 both ways = 33.9 %, stage 2 + pool + head = 66.1 %) / their summed CUDA-event durations
            tot_ms = sum(k["ms"] for k in kern.values())
            match_tflops = FLOP_PER_PAIR * T_loc * D / (tot_ms * 1e-3) / 1e12
            roof_kernel = " + ".join(sorted(kern)) + " (fused tcgen05 xcorr_eff)"
            # DRAM traffic per unit (pair, direction) from the ncu --set full capture in profiles/r01_ncu_pair_kernels.md
            # (dram__bytes_read.sum + dram__bytes_write.sum per launch / units per launch): p1a2 31.4 KB, p1b 49.8 KB, p2y 52.0 KB
            roof_traffic = 2 * T_loc * D * (31.4e3 + 49.8e3 + 52.0e3)
            roof_extra = {"kernels": {n: {"launches": k["launches"], "avg_ms_per_launch": k["ms"] / k["launches"],
                                          "share_of_match": k["ms"] / match_ms} for n, k in kern.items()}}
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.mode == "fast" else "f32",
            "data": "synthetic",
            "config": {"workload": f"configs[1]: Point Transformer encode of {T_loc} tracks/GPU + {D} detections x {NPTS} pts "
                                   f"(backbone_list {list(BLIST)}), {T_loc}x{D} all-pairs xcorr_eff match per GPU",
                       "mode": ("fast: fused bf16 tcgen05 matcher (pair_tc2.cu), fp32 accumulate/norms, tf32 tcgen05 SA shared MLPs, |dlogit| <= 3e-2"
                                if args.mode == "fast" else "parity: fp32 FFMA kernels, logits within 1e-4 of the reference"),
                       "l2": "no flush needed: each step streams >1 GB of activations (>> 126 MB L2)",
                       "sharding": f"track rows over {world} rank(s), one all-gather of detection embeddings"},
            "objects_encoded_per_s": n_enc * world / (enc_ms * 1e-3),
            "phase_ms": {"encode": enc_ms, "match": match_ms},
            "per_step_ms": per_step,
            "gpu_launches": launches,
            "clocks": clocks,
            "e2e": {"value": pairs_total / (e2e_ms / args.steps * 1e-3), "unit": "pairs/s",
                    "h2d_bytes_per_step": (tracks_h.numel() + dets_h.numel()) * 4, "d2h_bytes_per_step": out_h.numel() * 4},
            "roofline": {"bound": "tensor", "achieved": match_tflops, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": (match_tflops / peak_tf) if match_tflops else None, "traffic": roof_traffic if kern else None,
                         "per": "all fused launches of one step (three kernels x both directions x pair chunks), rank 0",
                         "kernel": roof_kernel, "peak_source": pk_src + " bf16 sustained (MEASURED_PEAKS.json)", **roof_extra},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_reference_sample(32, 2048)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
