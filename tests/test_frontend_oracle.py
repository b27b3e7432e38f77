"""CPU: the numpy restatement of the crop -> centre -> resample front-end (oracle/frontend_oracle.py) against closed-form
cases (the reference python itself needs mmcv / pytorch3d and cannot be imported here: parity unpinned, see the oracle header)."""
import numpy as np
import torch

from oracle import frontend_oracle as F


def scene(P, B, seed):
    g = np.random.default_rng(seed)
    pts = (g.standard_normal((P, 4)) * np.array([20.0, 20.0, 1.5, 1.0])).astype(np.float32)
    boxes = np.concatenate([g.uniform(-25, 25, (B, 2)), g.uniform(-1, 1, (B, 1)), g.uniform(1.5, 8.0, (B, 2)), g.uniform(1.0, 4.0, (B, 1)),
                            g.uniform(-np.pi, np.pi, (B, 1))], 1).astype(np.float32)
    return pts, boxes


def test_axis_aligned_boxes_are_plain_interval_tests():
    pts, boxes = scene(5000, 6, 0)
    boxes[:, 6] = 0.0
    inside, _ = F.points_in_boxes(boxes, pts)
    d = np.abs(pts[None, :, :3].astype(np.float64) - boxes[:, None, :3])
    ref = (d[..., 0] < boxes[:, None, 3] / 2) & (d[..., 1] < boxes[:, None, 4] / 2) & (d[..., 2] <= boxes[:, None, 5] / 2)
    assert (inside != ref).sum() <= 2          # only exact-boundary roundings may differ


def test_centred_coordinates_lie_in_the_box_and_invert_the_pose():
    pts, boxes = scene(20000, 8, 1)
    inside, _ = F.points_in_boxes(boxes, pts)
    lengths = inside.sum(1)
    assert lengths.max() > 10
    N = 64
    g = np.random.default_rng(2)
    rank = np.stack([g.integers(0, max(1, l), N) for l in lengths])
    out, ln = F.crop_center_resample(boxes, pts, N, rank)
    assert out.shape == (1, 8, N, 3) and torch.equal(ln[0], torch.from_numpy(lengths))
    o = out[0].numpy()
    for i in range(8):
        if lengths[i] == 0:
            assert (o[i] == 0).all()
            continue
        assert (np.abs(o[i, :, 0]) <= boxes[i, 3] / 2 + 1e-4).all() and (np.abs(o[i, :, 1]) <= boxes[i, 4] / 2 + 1e-4).all()
        assert (np.abs(o[i, :, 2]) <= boxes[i, 5] / 2 + 1e-4).all()
        # forward pose: p = Rz(-yaw) c + t  (get_affine_torch with rotation = -(0, 0, yaw))
        c, s = np.cos(boxes[i, 6]), np.sin(boxes[i, 6])
        back = np.stack([c * o[i, :, 0] + s * o[i, :, 1], -s * o[i, :, 0] + c * o[i, :, 1], o[i, :, 2]], 1) + boxes[i, :3]
        src = pts[inside[i]][rank[i], :3]
        assert np.abs(back - src).max() < 1e-4
