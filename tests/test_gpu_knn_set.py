"""-m gpu: unordered-set kNN (pcreid_knn_point_set) has exactly the members of the ordered torch-path kNN
(pcreid_knn_point, itself bit-exact against the oracle), including exact ties at the k-th boundary."""
import pytest
import torch

from oracle import reid_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _check(xyz, S, k):
    from pcreid_b200 import kernels as K
    new_xyz = xyz[:, :S].contiguous()
    ref = K.knn_point(k, xyz, new_xyz).sort(dim=2)[0]
    got = K.knn_point_set(k, xyz, new_xyz)
    assert got.shape == ref.shape and got.dtype == torch.int32
    assert torch.equal(got.sort(dim=2)[0], ref)


@pytest.mark.parametrize("N,S,k", [(256, 256, 32), (256, 128, 48), (128, 64, 48), (64, 32, 48), (160, 160, 32), (80, 40, 48),
                                   (1024, 1024, 32), (512, 256, 48), (40, 40, 40), (33, 7, 1),
                                   (700, 350, 48), (1000, 500, 32), (300, 300, 32), (1024, 512, 48)])
def test_members_match_ordered_knn(N, S, k):
    _check(O.synth_objects(7, N, 3).to(DEV).contiguous(), S, k)


@pytest.mark.parametrize("N,S,k", [(256, 256, 32), (128, 64, 48), (1024, 512, 32), (400, 200, 48)])
def test_ties_duplicated_points(N, S, k):
    g = torch.Generator().manual_seed(5)
    base = O.synth_objects(5, N, 4)
    U = max(2, N // 6)
    dup = torch.stack([base[b, :U][torch.randint(0, U, (N,), generator=g)] for b in range(5)])
    _check(dup.to(DEV).contiguous(), S, k)


def test_all_points_identical():
    xyz = torch.zeros(3, 128, 3, device=DEV)
    _check(xyz, 64, 48)
    _check(torch.full((2, 256, 3), 1.25, device=DEV), 256, 32)
    _check(torch.full((2, 1024, 3), -0.5, device=DEV), 512, 48)        # shared-memory-staged variant (N > 256)
