"""world_size-2 gloo test (CPU) of the row-sharded matcher: sharding arithmetic, the single all-gather of detection
embeddings and the optional score gather, with the C-ABI kernels replaced by their spec emulation."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, T, D, kind, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fake_kernels
    import helpers
    from oracle import reid_oracle as O
    from pcreid_b200.parallel import match_all_pairs_sharded, shard_range
    fake_kernels.install()
    torch.set_num_threads(2)
    m, orc = helpers.build_pair(kind)
    tracks, dets = O.synth_objects(T, 128, 0), O.synth_objects(D, 128, 1)
    t0, t1 = shard_range(T, rank, world)
    d0, d1 = shard_range(D, rank, world)
    tc = [shard_range(T, r, world)[1] - shard_range(T, r, world)[0] for r in range(world)]
    dc = [shard_range(D, r, world)[1] - shard_range(D, r, world)[0] for r in range(world)]
    rows = match_all_pairs_sharded(m, tracks[t0:t1], dets[d0:d1], dc)
    full = match_all_pairs_sharded(m, tracks[t0:t1], dets[d0:d1], dc, gather_scores=True, track_counts=tc)
    if rank == 0:
        xt, ht = orc.encode(tracks)
        xd, hd = orc.encode(dets)
        ref = orc.match_all_pairs(ht, xt, hd, xd)
        ret["rows_err"] = float((rows - ref[t0:t1]).abs().max())
        ret["full_err"] = float((full - ref).abs().max())
        ret["shape"] = tuple(full.shape)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_and_balances():
    sys.path.insert(0, ROOT)
    from pcreid_b200.parallel import shard_range
    for n, w in ((10, 3), (8, 8), (5, 8), (1024, 2), (0, 2)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
@pytest.mark.parametrize("kind", ["pt", "concat"])     # 'concat' all-gathers the pooled vectors only
def test_row_sharded_matcher_world2_gloo(kind):
    T, D = 5, 3      # uneven shards on purpose: tracks 3+2, detections 2+1
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000) + (7 if kind == "concat" else 0)
    mp.spawn(_worker, args=(2, port, T, D, kind, ret), nprocs=2, join=True)
    assert ret["shape"] == (T, D)
    assert ret["rows_err"] < 2e-5 and ret["full_err"] < 2e-5
