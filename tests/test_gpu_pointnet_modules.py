"""-m gpu: the mmdet3d PointNet++ module family (ops/pointnet_modules.py) against the CPU restatement
(oracle/pointnet_modules_oracle.py): sampled indices exact, features within 1e-4 (fp32, different summation order)."""
import pytest
import torch

from oracle import pointnet_modules_oracle as PO
from oracle import reid_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-4


def _randomise_bn(m, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.copy_(1 + 0.3 * torch.randn(mod.weight.shape, generator=g))
                mod.bias.copy_(0.2 * torch.randn(mod.bias.shape, generator=g))
                mod.running_mean.copy_(0.2 * torch.randn(mod.bias.shape, generator=g))
                mod.running_var.copy_(0.5 + torch.rand(mod.bias.shape, generator=g))
    return m


@pytest.mark.parametrize("normalize,dilated,use_feat", [(False, False, True), (True, True, True), (False, False, False)])
def test_point_sa_module_msg(normalize, dilated, use_feat):
    from pcreid_b200.ops import PointSAModuleMSG
    torch.manual_seed(3)
    cin = 6 if use_feat else 0
    m = _randomise_bn(PointSAModuleMSG(num_point=48, radii=[0.6, 1.2], sample_nums=[16, 32],
                                       mlp_channels=[[cin, 16, 32], [cin, 32, 32, 32]], normalize_xyz=normalize,
                                       dilated_group=dilated), 1).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    xyz = O.synth_objects(3, 200, 7)
    feat = torch.randn(3, cin, 200) if use_feat else None
    m = m.to(DEV)
    with torch.no_grad():
        nx, nf, idx = m(xyz.to(DEV), None if feat is None else feat.to(DEV))
    ox, of, oidx = PO.sa_module_msg(sd, 48, [0.6, 1.2], [16, 32], xyz, feat, normalize_xyz=normalize, dilated_group=dilated)
    assert torch.equal(idx.cpu().long(), oidx.long())
    assert torch.equal(nx.cpu(), ox)
    assert nf.shape == of.shape == (3, 64, 48)
    assert (nf.cpu() - of).abs().max() < TOL


def test_point_sa_module_fused_equal_width_and_group_all():
    from pcreid_b200.ops import PointSAModule, build_sa_module
    torch.manual_seed(4)
    m = _randomise_bn(build_sa_module(dict(type="PointSAModule", mlp_channels=[29, 32, 32, 32], num_point=40, radius=0.9,
                                           num_sample=24)), 2).eval()     # 29 + 3 xyz = 32 -> the fused three-layer kernel
    assert isinstance(m, PointSAModule)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    xyz, feat = O.synth_objects(2, 150, 8), torch.randn(2, 29, 150)
    with torch.no_grad():
        nx, nf, idx = m.to(DEV)(xyz.to(DEV), feat.to(DEV))
    ox, of, oidx = PO.sa_module_msg(sd, 40, [0.9], [24], xyz, feat)
    assert torch.equal(idx.cpu().long(), oidx.long()) and (nf.cpu() - of).abs().max() < TOL
    # given indices / GroupAll
    g = _randomise_bn(PointSAModule(mlp_channels=[29, 64, 128]), 3).eval()
    sdg = {k: v.clone() for k, v in g.state_dict().items()}
    with torch.no_grad():
        gx, gf, gi = g.to(DEV)(xyz.to(DEV), feat.to(DEV))
    _, ogf, _ = PO.sa_module_msg(sdg, None, [None], [None], xyz, feat)
    assert gx is None and gi is None and gf.shape == (2, 128, 1)
    assert (gf.cpu() - ogf).abs().max() < TOL


@pytest.mark.parametrize("with_target", [True, False])
def test_point_fp_module(with_target):
    from pcreid_b200.ops import PointFPModule
    torch.manual_seed(5)
    c1, c2 = (12 if with_target else 0), 20
    m = _randomise_bn(PointFPModule([c1 + c2, 48, 32]), 4).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    tgt, src = O.synth_objects(3, 130, 9), O.synth_objects(3, 40, 10)
    tf = torch.randn(3, c1, 130) if with_target else None
    sf = torch.randn(3, c2, 40)
    with torch.no_grad():
        got = m.to(DEV)(tgt.to(DEV), src.to(DEV), None if tf is None else tf.to(DEV), sf.to(DEV))
    ref = PO.fp_module(sd, tgt, src, tf, sf)
    assert got.shape == ref.shape == (3, 32, 130)
    assert (got.cpu() - ref).abs().max() < TOL


def test_training_mode_is_rejected():
    from pcreid_b200.ops import PointFPModule
    m = PointFPModule([8, 8]).to(DEV).train()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 4, 3, device=DEV), torch.zeros(1, 4, 3, device=DEV), None, torch.zeros(1, 8, 4, device=DEV))


# ---------------------------------------------------------------------------------------------------------------
# F-FPS / FS samplers and pool_mod='avg' (points_sampler.py:107-157, point_sa_module.py:144-164)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,M,C,norm", [(96, 96, 3, False), (130, 70, 35, False), (64, 200, 131, True), (257, 65, 16, True)])
def test_calc_square_dist_bit_exact_vs_oracle(N, M, C, norm):
    from oracle import ops_oracle as P
    from pcreid_b200.ops.furthest_point_sample import calc_square_dist
    g = torch.Generator().manual_seed(C)
    a, b = torch.randn(2, N, C, generator=g), torch.randn(2, M, C, generator=g)
    got = calc_square_dist(a.to(DEV), b.to(DEV), norm=norm).cpu()
    ref = P.pairwise_sqdist(a, b, norm=norm)
    assert torch.equal(got.isnan(), ref.isnan())
    ok = ~ref.isnan()                                   # norm=True takes sqrt of a rounded difference that may be < 0
    assert torch.equal(got[ok], ref[ok])
    assert (got[ok] - PO.calc_square_dist_ref(a, b, norm=norm)[ok]).abs().max() < 1e-4 * max(1.0, float(ref[ok].abs().max()))


@pytest.mark.parametrize("mods,ranges,npts", [(["F-FPS"], [-1], [32]), (["FS"], [-1], [24]), (["D-FPS", "F-FPS"], [64, -1], [16, 16])])
def test_points_sampler_ffps_fs(mods, ranges, npts):
    from pcreid_b200.ops import Points_Sampler
    xyz = O.synth_objects(3, 160, 11)
    feat = torch.randn(3, 32, 160, generator=torch.Generator().manual_seed(5))
    got = Points_Sampler(npts, mods, ranges)(xyz.to(DEV), feat.to(DEV)).cpu()
    assert got.dtype == torch.int32
    assert torch.equal(got, PO.points_sampler(xyz, feat, npts, mods, ranges))                                  # kernel arithmetic
    assert torch.equal(got, PO.points_sampler(xyz, feat, npts, mods, ranges, sqdist=PO.calc_square_dist_ref))  # reference formula


@pytest.mark.parametrize("fps_mod,equal_width", [(["D-FPS"], True), (["F-FPS"], False), (["FS"], False)])
def test_point_sa_module_avg_pool_and_feature_samplers(fps_mod, equal_width):
    from pcreid_b200.ops import PointSAModuleMSG
    torch.manual_seed(5)
    npt = 20 if fps_mod == ["FS"] else 40
    chans = [[29, 32, 32, 32], [29, 16, 48]] if equal_width else [[29, 24, 40], [29, 16, 48]]
    m = _randomise_bn(PointSAModuleMSG(num_point=npt, radii=[0.7, 1.3], sample_nums=[12, 24], mlp_channels=chans,
                                       fps_mod=fps_mod, fps_sample_range_list=[-1], pool_mod="avg"), 4).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    xyz, feat = O.synth_objects(2, 150, 9), torch.randn(2, 29, 150)
    with torch.no_grad():
        nx, nf, idx = m.to(DEV)(xyz.to(DEV), feat.to(DEV))
    ox, of, oidx = PO.sa_module_msg(sd, npt, [0.7, 1.3], [12, 24], xyz, feat, pool_mod="avg", fps_mod=fps_mod)
    S = 2 * npt if fps_mod == ["FS"] else npt
    assert idx.shape == (2, S) and torch.equal(idx.cpu().long(), oidx.long())
    assert torch.equal(nx.cpu(), ox)
    assert nf.shape == of.shape and (nf.cpu() - of).abs().max() < TOL
    # GroupAll with mean pooling
    from pcreid_b200.ops import PointSAModule
    g = _randomise_bn(PointSAModule(mlp_channels=[29, 64, 128], pool_mod="avg"), 3).eval()
    sdg = {k: v.clone() for k, v in g.state_dict().items()}
    with torch.no_grad():
        _, gf, _ = g.to(DEV)(xyz.to(DEV), feat.to(DEV))
    _, ogf, _ = PO.sa_module_msg(sdg, None, [None], [None], xyz, feat, pool_mod="avg")
    assert gf.shape == (2, 128, 1) and (gf.cpu() - ogf).abs().max() < TOL
