"""Row-sharded all-pairs matching over the GPUs of one box.

Objects are independent and pair (i, j) needs only objects i and j, so rank r encodes tracks
[r*T/G, (r+1)*T/G) and detections [r*D/G, (r+1)*D/G), ONE all-gather moves the detection embeddings
(per-point feature map + xyz for 'xcorr_eff', pooled vector for 'concat'), each rank scores its
T/G x D row block, and the row blocks are (optionally) gathered.  One process per GPU,
torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """contiguous, balanced split of range(n): the first n % world ranks get one extra element."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _all_gather_rows(t, counts, group):
    """all-gather of tensors that differ in dim 0 (row counts known on every rank)."""
    world = len(counts)
    if world == 1:
        return t
    mx = max(counts)
    pad = t
    if t.shape[0] < mx:
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
    out = torch.empty((world * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    if all(c == mx for c in counts):
        return out
    return torch.cat([out[r * mx:r * mx + c] for r, c in enumerate(counts)], dim=0)


def _encode_both(model, tracks_local, dets_local):
    """tracks and detections of equal point count go through the encoder as ONE batch (objects are independent; larger
    launches), exactly like ReIDNet.siamese_forward concatenates the two sides (ReIDNet.py:311-332)."""
    if tracks_local.shape[1:] == dets_local.shape[1:] and tracks_local.shape[0] > 0 and dets_local.shape[0] > 0:
        nt = tracks_local.shape[0]
        xyz, h = model.encode(torch.cat([tracks_local, dets_local], dim=0))
        return xyz[:nt], h[:nt], xyz[nt:], h[nt:]
    xyz_t, h_t = model.encode(tracks_local)
    xyz_d, h_d = model.encode(dets_local)
    return xyz_t, h_t, xyz_d, h_d


def encode_and_gather(model, tracks_local, dets_local, det_counts, group=None):
    """encode this rank's tracks / detection slice, all-gather the detection embeddings -> (xyz_t, h_t, xyz_d, h_d)."""
    xyz_t, h_t, xyz_d, h_d = _encode_both(model, tracks_local, dets_local)
    if len(det_counts) > 1:
        h_d = _all_gather_rows(h_d.contiguous(), det_counts, group)
        xyz_d = _all_gather_rows(xyz_d.contiguous(), det_counts, group)
    return xyz_t, h_t, xyz_d, h_d


class _RowGather:
    """all-gather of tensors that differ in dim 0, started asynchronously (NCCL runs it on its own stream) so that the
    caller can score the detections it already holds while the others are in flight."""

    def __init__(self, t, counts, group):
        self.counts, self.mx = counts, max(counts)
        pad = t
        if t.shape[0] < self.mx:
            pad = torch.zeros((self.mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            pad[:t.shape[0]] = t
        self.src = pad.contiguous()
        self.out = torch.empty((len(counts) * self.mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        self.work = dist.all_gather_into_tensor(self.out, self.src, group=group, async_op=True)

    def wait(self):
        self.work.wait()

    def block(self, r):
        """rows contributed by rank r (valid after wait())"""
        return self.out[r * self.mx:r * self.mx + self.counts[r]]


def match_all_pairs_sharded(model, tracks_local, dets_local, det_counts, pair_mask_rows=None, group=None,
                            gather_scores=False, track_counts=None, chunk=8192, overlap=True):
    """tracks_local (T_r, N, 3), dets_local (D_r, N, 3) on this rank; det_counts = [D_0 .. D_{G-1}].
    Returns this rank's (T_r, D) score rows, or the full (T, D) matrix on every rank if gather_scores
    (then track_counts = [T_0 .. T_{G-1}] is required).
    overlap: the all-gather of the detection embeddings runs while this rank scores its tracks against its OWN detection
    block; the blocks of the other ranks are scored as they become available."""
    world = len(det_counts)
    if getattr(model, "match_type", None) == 'concat':
        # the head only needs the pooled vectors: pool locally, all-gather 128 floats per detection instead of the maps
        with torch.no_grad():
            _, h_t, _, h_d = _encode_both(model, tracks_local, dets_local)
            e_t, e_d = model.pooled_embedding(h_t), model.pooled_embedding(h_d)
            if world > 1:
                e_d = _all_gather_rows(e_d.contiguous(), det_counts, group)
            rows = model.concat_all_pairs_pooled(e_t, e_d, pair_mask_rows)
    elif world > 1 and overlap:
        rank = dist.get_rank(group)
        xyz_t, h_t, xyz_d, h_d = _encode_both(model, tracks_local, dets_local)
        g_h, g_x = _RowGather(h_d.contiguous(), det_counts, group), _RowGather(xyz_d.contiguous(), det_counts, group)
        offs = [sum(det_counts[:r]) for r in range(world + 1)]
        rows = torch.zeros((h_t.shape[0], offs[-1]), device=h_t.device, dtype=torch.float32)

        def score(c0, c1, h_blk, xyz_blk):
            if c1 <= c0:
                return
            m = None if pair_mask_rows is None else pair_mask_rows[:, c0:c1].contiguous()
            rows[:, c0:c1] = model.match_all_pairs(h_t, xyz_t, h_blk, xyz_blk, pair_mask=m, chunk=chunk)

        score(offs[rank], offs[rank + 1], h_d, xyz_d)        # own block: needs nothing from the other ranks
        g_h.wait()
        g_x.wait()
        if min(det_counts) == max(det_counts):               # no padding: the gathered buffer is the detection list
            for c0, c1 in ((0, offs[rank]), (offs[rank + 1], offs[-1])):
                score(c0, c1, g_h.out[c0:c1], g_x.out[c0:c1])
        else:
            for r in range(world):
                if r != rank:
                    score(offs[r], offs[r + 1], g_h.block(r), g_x.block(r))
    else:
        xyz_t, h_t, xyz_d, h_d = encode_and_gather(model, tracks_local, dets_local, det_counts, group)
        rows = model.match_all_pairs(h_t, xyz_t, h_d, xyz_d, pair_mask=pair_mask_rows, chunk=chunk)
    if gather_scores and world > 1:
        return _all_gather_rows(rows.contiguous(), track_counts, group)
    return rows


def gathered_bytes(model, n_points, det_counts, track_counts=None, gather_scores=False, feat_channels=64):
    """bytes one rank RECEIVES per step from the two collectives of match_all_pairs_sharded."""
    D = sum(det_counts)
    per_det = 128 * 4 if getattr(model, "match_type", None) == 'concat' else (feat_channels + 3) * n_points * 4
    emb = per_det * D
    scores = 4 * sum(track_counts) * D if (gather_scores and track_counts) else 0
    return {"embeddings": emb, "scores": scores}
