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


def match_all_pairs_sharded(model, tracks_local, dets_local, det_counts, pair_mask_rows=None, group=None,
                            gather_scores=False, track_counts=None, chunk=8192):
    """tracks_local (T_r, N, 3), dets_local (D_r, N, 3) on this rank; det_counts = [D_0 .. D_{G-1}].
    Returns this rank's (T_r, D) score rows, or the full (T, D) matrix on every rank if gather_scores
    (then track_counts = [T_0 .. T_{G-1}] is required)."""
    if getattr(model, "match_type", None) == 'concat':
        # the head only needs the pooled vectors: pool locally, all-gather 128 floats per detection instead of the maps
        with torch.no_grad():
            _, h_t, _, h_d = _encode_both(model, tracks_local, dets_local)
            e_t, e_d = model.pooled_embedding(h_t), model.pooled_embedding(h_d)
            if len(det_counts) > 1:
                e_d = _all_gather_rows(e_d.contiguous(), det_counts, group)
            rows = model.concat_all_pairs_pooled(e_t, e_d, pair_mask_rows)
    else:
        xyz_t, h_t, xyz_d, h_d = encode_and_gather(model, tracks_local, dets_local, det_counts, group)
        rows = model.match_all_pairs(h_t, xyz_t, h_d, xyz_d, pair_mask=pair_mask_rows, chunk=chunk)
    if gather_scores and len(det_counts) > 1:
        return _all_gather_rows(rows.contiguous(), track_counts, group)
    return rows
