"""-m gpu: the torch custom-op layer on the device -- torch.library.opcheck (schema / mutation annotations, fake kernel,
autograd registration, AOT dispatch) on representative ops, and QueryAndGroup (one fused kernel behind
torch.ops.pcreid.query_group) bit-exact against the oracle that is pinned to the reference's own class."""
import pytest
import torch

import pcreid_b200.ops  # noqa: F401  (registers torch.ops.pcreid.*)
from oracle import ops_oracle as P
from oracle import pointnet_modules_oracle as PO
from oracle import reid_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _opcheck(op, args):
    torch.library.opcheck(op, args, test_utils=("test_schema", "test_faketensor", "test_autograd_registration", "test_aot_dispatch_dynamic"))


def test_opcheck_point_ops():
    x = O.synth_objects(2, 128, 3).to(DEV)
    c = x[:, :16].contiguous()
    ops = torch.ops.pcreid
    _opcheck(ops.fps.default, (2, 128, 16, x, torch.full((2, 128), 1e10, device=DEV), torch.empty((2, 16), dtype=torch.int32, device=DEV)))
    _opcheck(ops.knn_t.default, (2, 128, 16, 8, x, c, torch.zeros((2, 8, 16), dtype=torch.int32, device=DEV), torch.zeros((2, 8, 16), device=DEV)))
    _opcheck(ops.ball_query.default, (2, 128, 16, 0.0, 0.8, 8, c, x, torch.zeros((2, 16, 8), dtype=torch.int32, device=DEV)))
    f = torch.randn(2, 12, 128, device=DEV)
    idx = torch.randint(0, 128, (2, 16, 8), dtype=torch.int32, device=DEV)
    _opcheck(ops.group_points.default, (2, 12, 128, 16, 8, f, idx, torch.empty((2, 12, 16, 8), device=DEV)))
    _opcheck(ops.gather_points.default, (2, 12, 128, 16, f, idx[:, :, 0].contiguous(), torch.empty((2, 12, 16), device=DEV)))
    _opcheck(ops.query_group.default, (2, 12, 128, 16, 8, x, c, f, idx, 1, 0.0, torch.empty((2, 15, 16, 8), device=DEV), None))
    _opcheck(ops.knn_point.default, (2, 128, 16, 8, x, c, torch.empty((2, 16, 8), dtype=torch.int32, device=DEV)))


def test_opcheck_struct_and_fused_ops():
    from pcreid_b200 import torch_ops as T
    x = torch.randn(2, 16, 64, device=DEV)
    w = torch.randn(16, 32, device=DEV)
    y = torch.empty(2, 32, 64, device=DEV)
    a = T.linear_args()
    a.B, a.rows, a.CO, a.K1 = 2, 64, 32, 16
    a.X1, a.x1_bs, a.ldx1, a.W1 = x, x.stride(0), x.stride(1), w
    a.Y, a.y_bs, a.ldy = y, y.stride(0), y.stride(1)
    _opcheck(torch.ops.pcreid.cn_linear.default, a.astuple())       # opcheck works on clones of the arguments
    torch.ops.pcreid.cn_linear(*a.astuple())
    assert torch.allclose(y, torch.einsum("bkn,kc->bcn", x, w), atol=1e-4)
    rc = torch.ops.pcreid.cn_linear_tc(*a.astuple())          # K = 16 is outside the tensor-core kernel's tiles: answers UNSUPPORTED
    assert rc in (0, 3)


def test_ops_use_the_current_stream_and_device_of_their_tensors():
    x = O.synth_objects(2, 128, 4).to(DEV)
    from pcreid_b200.ops import furthest_point_sample
    ref = P.furthest_point_sample(x.cpu(), 32)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        got = furthest_point_sample(x, 32)
    s.synchronize()
    assert torch.equal(got.cpu(), ref)


@pytest.mark.parametrize("max_r,normalize,uniform,use_xyz,use_feat", [
    (None, False, False, True, True), (0.9, True, False, True, True), (0.9, False, True, True, True), (0.7, False, True, True, False),
    (0.9, False, False, False, True), (0.9, True, False, True, False)])
def test_query_and_group_fused_bit_exact(max_r, normalize, uniform, use_xyz, use_feat):
    from pcreid_b200.ops import QueryAndGroup
    x = O.synth_objects(3, 160, 6)
    f = torch.randn(3, 21, 160, generator=torch.Generator().manual_seed(1)) if use_feat else None
    centers = x[:, :20].contiguous()
    qg = QueryAndGroup(max_r, 8, use_xyz=use_xyz, normalize_xyz=normalize, uniform_sample=uniform, return_grouped_xyz=True,
                       return_unique_cnt=uniform, return_grouped_idx=True)
    torch.manual_seed(11)
    got = qg(x.to(DEV), centers.to(DEV), None if f is None else f.to(DEV))
    torch.manual_seed(11)
    nf, gx, cnt, idx = PO.query_and_group_full(x, centers, f, max_r, 8, use_xyz=use_xyz, normalize_xyz=normalize, uniform_sample=uniform)
    assert torch.equal(got[-1].cpu(), idx)
    assert torch.equal(got[0].cpu(), nf) and torch.equal(got[1].cpu(), gx)
    if uniform:
        assert torch.equal(got[2], cnt)
    plain = QueryAndGroup(max_r, 8, use_xyz=use_xyz, normalize_xyz=normalize)
    if not uniform:
        assert torch.equal(plain(x.to(DEV), centers.to(DEV), None if f is None else f.to(DEV)).cpu(), nf)
