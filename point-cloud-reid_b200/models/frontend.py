"""Crop -> centre -> resample front-end: one LiDAR sweep + boxes -> the (B, N, 3) encoder input.

Mirrors the deprecated tracker's ``interpolate_per_frame`` + ``get_input_batch``
(mmdet3d/models/trackers/deprecated/pc_utils.py:31-96): points inside each box (DepthInstance3DBoxes with origin
(0.5, 0.5, 0.5), core/bbox/structures/depth_box3d.py:256-282), expressed in the box frame (inverse of the box pose),
resampled WITH replacement to ``subsample_number`` points; boxes without points give zeros.  The reference builds a (P, B)
mask, B python crops, a padded (B, Lmax, 3) batch and its homogeneous copy; here two kernels (csrc/frontend.cu) go from
the sweep to the encoder input through a bit mask."""
import torch

from .. import _lib
from .. import torch_ops as _T

_OPS = _T.ops


def points_in_boxes_mask(bboxes, pts):
    """-> (mask (B, ntiles, 32) int32 bit mask, counts (B, ntiles) int32, lengths (B,) int64)."""
    if not (bboxes.is_cuda and pts.is_cuda):
        raise RuntimeError("pcreid_b200 kernels need CUDA tensors (there is no CPU fallback)")
    assert bboxes.dtype == torch.float32 and pts.dtype == torch.float32 and bboxes.shape[1] == 7
    bboxes, pts = bboxes.contiguous(), pts.contiguous()
    P, B = pts.shape[0], bboxes.shape[0]
    L = _lib.lib()
    nt = L.pcreid_crop_tiles(P)
    mask = torch.empty((B, nt, 32), device=pts.device, dtype=torch.int32)
    counts = torch.empty((B, nt), device=pts.device, dtype=torch.int32)
    _OPS.crop_mask(P, B, pts, pts.shape[1], bboxes, mask, counts)
    return mask, counts, counts.sum(1, dtype=torch.int64)


def crop_center_resample(bboxes, pts, subsample_number, sample_rank=None, generator=None):
    """bboxes (B, 7) = (x, y, z centre, dx, dy, dz, yaw), pts (P, >=3) -> (out (1, B, N, 3), lengths (1, B)).

    ``sample_rank`` (B, N) int64: which in-box point (in point order) each output slot takes -- what the reference draws with
    ``torch.randint(high=length)`` per box (pc_utils.py:84-85).  When omitted, ranks are drawn on the device as
    floor(U[0,1) * length): the same distribution, not the reference's host random stream."""
    P, B, N = pts.shape[0], bboxes.shape[0], int(subsample_number)
    if P == 0 or B == 0:        # empty sweep / no boxes: zero crops with zero lengths, as the reference returns (pc_utils.py:88-92)
        if not (bboxes.is_cuda and pts.is_cuda):
            raise RuntimeError("pcreid_b200 kernels need CUDA tensors (there is no CPU fallback)")
        return (torch.zeros((1, B, N, 3), device=pts.device, dtype=torch.float32),
                torch.zeros((1, B), device=pts.device, dtype=torch.int64))
    mask, counts, lengths = points_in_boxes_mask(bboxes, pts)
    bboxes, pts = bboxes.contiguous(), pts.contiguous()
    prefix = torch.zeros((B, counts.shape[1] + 1), device=pts.device, dtype=torch.int32)
    prefix[:, 1:] = torch.cumsum(counts, 1, dtype=torch.int32)
    if sample_rank is None:
        u = torch.rand((B, N), device=pts.device, generator=generator)
        sample_rank = (u * lengths[:, None]).long()
    sample_rank = sample_rank.to(device=pts.device, dtype=torch.int64).contiguous()
    assert sample_rank.shape == (B, N)
    out = torch.empty((B, N, 3), device=pts.device, dtype=torch.float32)
    _OPS.crop_gather(P, B, N, pts, pts.shape[1], bboxes, mask, prefix, sample_rank, out)
    return out.unsqueeze(0), lengths.unsqueeze(0)
