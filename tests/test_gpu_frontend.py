"""-m gpu: crop -> centre -> resample front-end (csrc/frontend.cu) against the numpy restatement of the tracker's
interpolate_per_frame + get_input_batch, and its in-box test bit-for-bit against the reference's own points_in_boxes_cuda.cu
(compiled unmodified into oracle/_ref/libref_pib.so when the image's torch headers allow)."""
import numpy as np
import pytest
import torch

from oracle import frontend_oracle as F
from test_frontend_oracle import scene

pytestmark = pytest.mark.gpu
DEV = "cuda"


def decode(mask, P):
    """(B, ntiles, 32) int32 bit mask -> bool (B, P)."""
    m = mask.cpu().numpy().astype(np.uint32).reshape(mask.shape[0], -1)
    bits = ((m[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool).reshape(mask.shape[0], -1)
    return bits[:, :P]


@pytest.mark.parametrize("P,B", [(5000, 7), (1024, 33), (70001, 64), (17, 1)])
def test_mask_matches_oracle_and_reference_kernel(P, B):
    from pcreid_b200.models.frontend import points_in_boxes_mask
    pts, boxes = scene(P, B, 3)
    mask, counts, lengths = points_in_boxes_mask(torch.from_numpy(boxes).to(DEV), torch.from_numpy(pts).to(DEV))
    got = decode(mask, P)
    inside, margin = F.points_in_boxes(boxes, pts)
    diff = got != inside
    assert (margin[diff] < 1e-5).all(), "in-box decisions may only differ within rounding of a box face (cosf/sinf last bit)"
    assert diff.sum() <= 3
    assert torch.equal(lengths.cpu(), torch.from_numpy(got.sum(1)))
    assert int(counts.sum()) == int(got.sum())
    if F.ref_pib_available():
        b = torch.from_numpy(boxes).to(DEV)
        p = torch.from_numpy(pts[:, :3]).to(DEV)
        # what DepthInstance3DBoxes(origin=(.5,.5,.5)).points_in_boxes feeds the kernel (depth_box3d.py:270-280)
        bl = torch.stack([b[:, 1], -b[:, 0], b[:, 2] + b[:, 5] * -0.5, b[:, 4], b[:, 3], b[:, 5], b[:, 6]], 1).unsqueeze(0).contiguous()
        pl = torch.stack([p[:, 1], -p[:, 0], p[:, 2]], 1).unsqueeze(0).contiguous()
        ref = F.ref_points_in_boxes_lidar(bl, pl)[0].bool().t().cpu().numpy()          # (B, P)
        assert (got == ref).all(), "bit mask differs from the reference points_in_boxes kernel"


def test_crop_center_resample_matches_oracle():
    from pcreid_b200.models.frontend import crop_center_resample, points_in_boxes_mask
    P, B, N = 30000, 12, 128
    pts, boxes = scene(P, B, 4)
    boxes[3, :3] = 500.0                                                     # an empty box -> zeros
    bt, pt = torch.from_numpy(boxes).to(DEV), torch.from_numpy(pts).to(DEV)
    mask, _, lengths = points_in_boxes_mask(bt, pt)
    inside = decode(mask, P)
    g = np.random.default_rng(5)
    rank = np.stack([g.integers(0, max(1, l), N) for l in inside.sum(1)])
    out, ln = crop_center_resample(bt, pt, N, sample_rank=torch.from_numpy(rank))
    oo, lo = F.crop_center_resample(boxes, pts, N, rank, inside=inside)
    assert out.shape == (1, B, N, 3) and torch.equal(ln.cpu(), lo)
    assert (out.cpu() - oo).abs().max() < 1e-4                               # tolerance: the reference inverts a 4x4 pose matrix
    assert (out[0, 3] == 0).all() and int(ln[0, 3]) == 0


def test_device_side_ranks_and_scale_properties():
    """a full sweep (250k points, 300 boxes): every resampled point lies inside its box, lengths are the mask popcounts"""
    from pcreid_b200.models.frontend import crop_center_resample
    pts, boxes = scene(250000, 300, 6)
    bt, pt = torch.from_numpy(boxes).to(DEV), torch.from_numpy(pts).to(DEV)
    out, ln = crop_center_resample(bt, pt, 256, generator=torch.Generator(device=DEV).manual_seed(0))
    o = out[0]
    half = bt[:, None, 3:6] / 2 + 1e-4
    assert (o.abs() <= half).all()
    empty = ln[0] == 0
    assert (o[empty] == 0).all()
    inside, _ = F.points_in_boxes(boxes, pts)
    assert (ln[0].cpu() - torch.from_numpy(inside.sum(1))).abs().max() <= 2
    # feeds the encoder unchanged: (B, N, 3) float32 contiguous
    assert o.dtype == torch.float32 and o.is_contiguous()


def test_reidentifier_from_sweep_equals_manual_pipeline():
    """sweep -> crop / centre / resample -> encode -> class gate -> cost matrix in one call == the same steps by hand"""
    import helpers
    from oracle import reid_oracle as O
    from pcreid_b200.models.frontend import crop_center_resample
    from pcreid_b200.models.tracking import PointFeatureSet, PointReidentifier
    m, orc = helpers.build_pair("pt", (128, 64, 32), device=DEV)
    pts, boxes = scene(60000, 9, 8)
    bt, pt = torch.from_numpy(boxes).to(DEV), torch.from_numpy(pts).to(DEV)
    rank = torch.randint(0, 1 << 30, (9, 128), generator=torch.Generator().manual_seed(1))
    bank = PointFeatureSet()
    tr = O.synth_objects(5, 128, 9).to(DEV)
    xt, ht = m.encode(tr)
    bank.store_new(xt, ht, torch.full((5,), 128, device=DEV))
    reid = PointReidentifier(m, bank, subsample_number=128)
    dl = torch.tensor([0, 1, 2, 0, 1, 2, 0, 1, 9], device=DEV)
    tl = torch.tensor([0, 1, 2, 0, 7], device=DEV)
    ti = torch.arange(5, device=DEV)
    cost, xyz_d, h_d, len_d = reid.from_sweep(pt, bt, dl, ti, tl, sample_rank=rank)
    batch, lengths = crop_center_resample(bt, pt, 128, sample_rank=rank)
    cost2, _, _ = reid(batch[0], dl, lengths[0], ti, tl)
    assert torch.equal(cost, cost2) and torch.equal(len_d, lengths[0]) and cost.shape == (5, 9)
    # against the oracle on the same crops
    oxd, ohd = orc.encode(batch[0].cpu())
    mask = O.class_gated_pairs(tl.cpu(), torch.full((5,), 128), dl.cpu(), lengths[0].cpu())
    Lo = orc.match_all_pairs(ht.cpu(), xt.cpu(), ohd, oxd, pair_mask=mask)
    assert (cost.cpu() - Lo).abs().max() < 1e-4
    assert (cost[:, 8] == 0).all()                                           # class 9 is never compared
