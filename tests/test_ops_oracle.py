"""CPU checks of the point-op oracle (oracle/ops_oracle.c): known-answer properties the survey measured on the
reference kernels, and the canonical distance arithmetic the CUDA kernels reproduce."""
import numpy as np
import pytest
import torch

from oracle import ops_oracle as P
from oracle import reid_oracle as O


def _bitrev(v, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (v & 1)
        v >>= 1
    return r


def fps_keyrule(xyz, m):
    """Restates the CUDA kernel's selection rule (csrc/ops.cu fps_kernel): arg-max of min-distance with ties broken
    by the smallest (bitreverse(k mod bs), k) -- the closed form of the reference's shared-memory tree."""
    B, N, _ = xyz.shape
    bs = P.fps_block_size(N)
    bits = bs.bit_length() - 1
    x = xyz.numpy().astype(np.float32)
    out = np.zeros((B, m), np.int32)
    tie = np.array([(_bitrev(k % bs, bits) << 32) | k for k in range(N)], dtype=np.int64)
    for b in range(B):
        temp = np.full(N, 1e10, np.float32)
        old = 0
        for j in range(1, m):
            d = (x[b] - x[b, old]).astype(np.float32)
            t = (d[:, 1] * d[:, 1]).astype(np.float32)
            t = np.float32(d[:, 0].astype(np.float64) * d[:, 0].astype(np.float64) + t.astype(np.float64)).astype(np.float32)
            t = np.float32(d[:, 2].astype(np.float64) * d[:, 2].astype(np.float64) + t.astype(np.float64)).astype(np.float32)
            temp = np.minimum(t, temp)
            best = temp.max()
            cand = np.nonzero(temp == best)[0]
            old = int(cand[np.argmin(tie[cand])])
            out[b, j] = old
    return torch.from_numpy(out)


@pytest.mark.parametrize("N,M,dup", [(256, 64, False), (160, 48, True), (96, 96, True), (1000, 32, True), (7, 7, False)])
def test_fps_tie_rule_closed_form(N, M, dup):
    x = O.synth_objects(3, N, 11, dup=dup)
    assert torch.equal(P.furthest_point_sample(x, M), fps_keyrule(x, M))


def test_fps_degenerate_cloud_returns_zero():
    x = torch.zeros(2, 64, 3)
    assert (P.furthest_point_sample(x, 16) == 0).all()


def test_knn_heap_equals_canonical_without_ties():
    x = O.synth_objects(2, 200, 3)
    idx, d2 = P.knn(16, x, x[:, :50], return_dist=True)
    dd = ((x[:, :50, None, :] - x[:, None, :, :]) ** 2).sum(-1)
    ref = torch.sort(dd, dim=-1, stable=True)[1][..., :16]
    assert torch.equal(idx.transpose(1, 2).long(), ref)
    assert (d2[..., 1:] >= d2[..., :-1]).all()


def test_knn_fewer_points_than_k_pads_with_sentinel():
    x = O.synth_objects(1, 5, 4)
    idx, d2 = P.knn(8, x, x, return_dist=True)
    assert (idx[:, 5:, :] == 0).all() and (d2[..., 5:] == 1e10).all()


def test_ball_query_semantics():
    x = torch.tensor([[[0., 0, 0], [0.5, 0, 0], [2, 0, 0], [0.1, 0, 0], [0, 0, 0]]])
    q = torch.tensor([[[0., 0, 0], [10, 0, 0]]])
    idx = P.ball_query(0.2, 1.0, 4, x, q)
    assert idx[0, 0].tolist() == [0, 1, 4, 0]      # d2==0 always accepted; first hit pads the tail
    assert idx[0, 1].tolist() == [0, 0, 0, 0]      # no hit: row stays zero


def test_expansion_form_distance_is_bit_exact_vs_torch():
    """torch CPU square_distance == fma-chain + (x*x+y*y)+z*z + two adds."""
    for seed, dup in ((0, False), (1, True)):
        x = O.synth_objects(3, 256, seed, dup=dup)
        assert torch.equal(P.sqdist_expand(x[:, :128], x), O.square_distance(x[:, :128], x))


@pytest.mark.parametrize("C", [3, 64, 128])
def test_dgcnn_pairwise_is_bit_exact_vs_torch(C):
    x = torch.randn(2, C, 160, generator=torch.Generator().manual_seed(C))
    assert torch.equal(P.dgcnn_pd(x), O.dgcnn_pairwise(x))


def test_group_gather():
    f = torch.randn(2, 5, 9)
    idx = torch.randint(0, 9, (2, 4, 3), dtype=torch.int32)
    g = P.grouping_operation(f, idx)
    assert torch.equal(g, torch.gather(f.unsqueeze(2).expand(2, 5, 4, 9), 3, idx.long().unsqueeze(1).expand(2, 5, 4, 3)))
    i2 = torch.randint(0, 9, (2, 6), dtype=torch.int32)
    assert torch.equal(P.gather_points(f, i2), torch.gather(f, 2, i2.long().unsqueeze(1).expand(2, 5, 6)))


def test_three_nn_oracle_semantics():
    """three_nn_cuda.cu:11-66: ascending (d2, index) on ties, sentinel (0, +inf) when fewer than 3 sources exist."""
    t, s = O.synth_objects(2, 50, 0), O.synth_objects(2, 40, 1)
    d, idx = P.three_nn(t, s)
    d2 = ((t[:, :, None, :] - s[:, None, :, :]) ** 2).sum(-1)
    ref = torch.sort(d2, dim=2, stable=True)[1][:, :, :3]
    assert torch.equal(idx.long(), ref)                       # tie-free input: plain 3 nearest
    assert torch.allclose(d, torch.gather(d2, 2, ref).sqrt(), atol=1e-6)
    dup = O.synth_objects(1, 30, 2, dup=True)                 # duplicate points: lower index first
    _, idup = P.three_nn(dup, dup)
    d2d = ((dup[:, :, None, :] - dup[:, None, :, :]) ** 2).sum(-1)
    assert torch.equal(idup.long(), torch.sort(d2d, dim=2, stable=True)[1][:, :, :3])
    d1, i1 = P.three_nn(t, s[:, :2].contiguous())             # M = 2 < 3
    assert (i1[:, :, 2] == 0).all() and torch.isinf(d1[:, :, 2]).all()


def test_three_interpolate_oracle():
    f = torch.randn(2, 5, 20, generator=torch.Generator().manual_seed(0))
    idx = torch.randint(0, 20, (2, 9, 3), generator=torch.Generator().manual_seed(1)).int()
    w = torch.rand(2, 9, 3, generator=torch.Generator().manual_seed(2))
    out = P.three_interpolate(f, idx, w)
    ref = (torch.gather(f[:, :, None, :].expand(2, 5, 9, 20), 3, idx.long()[:, None].expand(2, 5, 9, 3)) * w[:, None]).sum(-1)
    assert torch.allclose(out, ref, atol=1e-6)


# ---------------------------------------------------------------------------------------------------------------
# F-FPS / FS samplers (points_sampler.py:107-157): the kernel-arithmetic distance matrix against the reference's torch formula
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,M,C", [(96, 96, 3), (130, 70, 35), (64, 200, 131)])
def test_pairwise_sqdist_oracle_vs_reference_formula(N, M, C):
    from oracle import pointnet_modules_oracle as PO
    g = torch.Generator().manual_seed(C)
    a, b = torch.randn(2, N, C, generator=g), torch.randn(2, M, C, generator=g)
    d = P.pairwise_sqdist(a, b, norm=False)
    ref = PO.calc_square_dist_ref(a, b, norm=False)
    assert d.shape == (2, N, M)
    assert (d - ref).abs().max() <= 2e-5 * ref.abs().max()               # same quantity, different summation order
    dn = P.pairwise_sqdist(a, b, norm=True)
    assert (dn - PO.calc_square_dist_ref(a, b, norm=True)).abs().max() < 1e-5
    # exact properties of the fma-chain arithmetic: zero diagonal and symmetry of the self-distance matrix
    s = P.pairwise_sqdist(a, a, norm=False)
    assert torch.equal(s, s.transpose(1, 2)) and (torch.diagonal(s, dim1=1, dim2=2) == 0).all()


@pytest.mark.parametrize("mods,ranges,npts", [(["F-FPS"], [-1], [32]), (["FS"], [-1], [24]), (["D-FPS", "F-FPS"], [64, -1], [16, 16])])
def test_samplers_same_indices_with_kernel_and_reference_distance(mods, ranges, npts):
    """F-FPS picks arg-maxes of the distance matrix: the fma-chain matrix and the reference's matmul matrix differ in the last
    bits only and select the same points on the seeded (continuous) inputs."""
    from oracle import pointnet_modules_oracle as PO
    xyz = O.synth_objects(3, 160, 11)
    feat = torch.randn(3, 32, 160, generator=torch.Generator().manual_seed(5))
    got = PO.points_sampler(xyz, feat, npts, mods, ranges)
    ref = PO.points_sampler(xyz, feat, npts, mods, ranges, sqdist=PO.calc_square_dist_ref)
    assert got.shape == (3, sum(n * (2 if m == "FS" else 1) for n, m in zip(npts, mods)))
    assert torch.equal(got, ref)
