"""BASELINE configs[4]: 10 Hz tracking replay sweep -- N tracks x N detections (256 pts, Point Transformer, xcorr_eff) from 256 x 256
to 16384 x 16384 pairs, a FIXED matrix per size sharded by track rows over the ranks (strong scaling), timed region = encode of
this rank's shard + all-gather of the detection embeddings + scoring of the row block + all-gather of the score rows.
  python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 scripts/bench_sweep.py [--sizes 256,1024,4096,16384]
One JSON line per size (rank 0), ms = max over ranks (CUDA events, barrier + synchronize on both sides)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pcreid_b200 import synthetic as S  # noqa: E402
from pcreid_b200.models import build_model  # noqa: E402
from pcreid_b200.parallel import gathered_bytes, match_all_pairs_sharded, shard_range  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="256,1024,4096")
ap.add_argument("--mode", default="parity_tc")
ap.add_argument("--graphs", type=int, default=1)
ap.add_argument("--max-seconds", type=float, default=40.0, help="per size: steps are chosen so that the timed region stays below this")
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(66)
model = build_model(S.point_transformer_cfg((256, 128, 64))).eval().to(dev)
model.set_mode(args.mode)
model.enable_cuda_graphs(bool(args.graphs))      # encode per shape; the match of per-frame sized blocks (T*D <= 16384 per call)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for T in [int(v) for v in args.sizes.split(",")]:
    D = T
    t0, t1 = shard_range(T, rank, world)
    d0, d1 = shard_range(D, rank, world)
    tc = [shard_range(T, r, world)[1] - shard_range(T, r, world)[0] for r in range(world)]
    dc = [shard_range(D, r, world)[1] - shard_range(D, r, world)[0] for r in range(world)]
    tracks = S.synth_objects(T, 256, 0)[t0:t1].contiguous().to(dev)
    dets = S.synth_objects(D, 256, 1)[d0:d1].contiguous().to(dev)
    step = lambda: match_all_pairs_sharded(model, tracks, dets, dc, gather_scores=True, track_counts=tc)
    est = T * D / (6.5e6 * world)                         # seconds per step at the single-GPU rate
    warm = 3 if est < 1.0 else 1
    steps = max(1, min(20, int(args.max_seconds / max(est, 1e-3)) - warm))
    for _ in range(warm):
        out = step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); e0.record()
    for _ in range(steps):
        out = step()
    e1.record(); barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    assert out.shape == (T, D) and bool(torch.isfinite(out).all())
    if rank == 0:
        print(json.dumps({"workload": f"configs[4] sweep: PT encode of {T}+{D} objects x 256 pts + {T}x{D} all-pairs xcorr_eff, row-sharded",
                          "n_gpus": world, "mode": args.mode, "cuda_graphs": bool(args.graphs), "scaling": "strong", "tracks": T, "dets": D, "steps": steps, "warmup": warm,
                          "ms_per_step_max_over_ranks": ms, "pairs_per_s": T * D / (ms * 1e-3), "objects_per_s": (T + D) / (ms * 1e-3),
                          "bytes_received_per_rank": gathered_bytes(model, 256, dc, tc, gather_scores=True)}), flush=True)
    del tracks, dets, out
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
