"""Host-side logic of the product modules (weight packing, eval-BatchNorm folding, first-conv
factorisation, EdgeConv factorisation, object index maps, layouts) checked on CPU against the oracle by
substituting the C-ABI kernels with a torch emulation of their *specification* (tests/fake_kernels.py)."""
import pytest
import torch

import fake_kernels
import helpers
from oracle import reid_oracle as O


@pytest.fixture()
def fake(monkeypatch):
    fake_kernels.install(monkeypatch)


@pytest.mark.parametrize("kind", ["pt", "concat", "dgcnn", "pointnet", "pt15m", "pt7m", "xcorr", "xcorr-baseline"])
def test_model_vs_oracle_with_spec_kernels(fake, kind):
    m, orc = helpers.build_pair(kind)
    t, d = O.synth_objects(3, 128, 0), O.synth_objects(4, 128, 1)
    with torch.no_grad():
        xt, ht = m._encode(t)
        xd, hd = m._encode(d)
    oxt, oht = orc.encode(t)
    oxd, ohd = orc.encode(d)
    assert (ht - oht).abs().max() < 2e-5 and (hd - ohd).abs().max() < 2e-5
    L = m.match_all_pairs(ht, xt, hd, xd, chunk=5)
    Lo = orc.match_all_pairs(oht, oxt, ohd, oxd)
    assert (L - Lo).abs().max() < 2e-5
    mask = torch.rand(3, 4, generator=torch.Generator().manual_seed(0)) > 0.4
    Lm = m.match_all_pairs(ht, xt, hd, xd, pair_mask=mask, chunk=5)
    assert (Lm - orc.match_all_pairs(oht, oxt, ohd, oxd, pair_mask=mask)).abs().max() < 2e-5
    assert (Lm[~mask] == 0).all()


def test_reference_api_surface(fake):
    m, orc = helpers.build_pair("pt")
    s1, s2 = O.synth_objects(3, 128, 2), O.synth_objects(3, 128, 3)
    x1, x2, h1, h2 = m.siamese_forward(s1, s2)
    assert h1.shape == (3, 64, 128) and x1.shape == (3, 128, 3)
    lg = m.match_forward_inference(h1, h2, x1, x2)
    o = orc.siamese_forward(s1, s2)
    assert (lg - orc.match_forward_inference(o[2], o[3], o[0], o[1])).abs().max() < 2e-5
    out, o1, o2 = m.xcorr_eff(h1, x1, h2, x2)
    assert out.shape == (3, 64, 256)
    assert m.get_pooled_feats(out).shape == (3, 128)
    ids = [torch.tensor([i]) for i in range(3)]
    lab = [torch.tensor([1]) for _ in range(3)]
    res = m(return_loss=False, sparse_1=list(s1), sparse_2=list(s2), dense_1=list(s1), dense_2=list(s2), label_1=lab, label_2=lab,
            id_1=ids, id_2=ids, size_1=lab, size_2=lab, vis_1=lab, vis_2=lab)
    assert isinstance(res, list) and res[0]['val_match_preds'].shape == (3,) and res[0]['val_match_gt'].tolist() == [1., 1., 1.]
    with pytest.raises(NotImplementedError):
        m(return_loss=True)


def test_forward_test_refuses_configs_it_does_not_evaluate(fake):
    """auxiliary heads that are built and enabled, the dense loss, a disabled match loss: forward_test raises instead of
    silently returning zeros / None where the reference would have run them (ReIDNet.py:652-662)."""
    m, _ = helpers.build_pair("pt")
    s1, s2 = O.synth_objects(2, 128, 2), O.synth_objects(2, 128, 3)
    one = [torch.tensor([1]) for _ in range(2)]
    kw = dict(sparse_1=list(s1), sparse_2=list(s2), dense_1=list(s1), dense_2=list(s2), label_1=one, label_2=one, id_1=one, id_2=one,
              size_1=one, size_2=one, vis_1=one, vis_2=one)
    m.losses_to_use.update(kl=False, cls=False, fp=False, shape=False, dense=False, match=True)
    assert m(return_loss=False, **kw)[0]['val_kl_loss'].item() == 0.
    m.losses_to_use["dense"] = True
    with pytest.raises(NotImplementedError):
        m(return_loss=False, **kw)
    m.losses_to_use.update(dense=False, cls=True)
    m.cls_head = torch.nn.Linear(128, 10)
    with pytest.raises(NotImplementedError):
        m(return_loss=False, **kw)
    m.cls_head = None
    m.losses_to_use["match"] = False
    with pytest.raises(NotImplementedError):
        m(return_loss=False, **kw)


def test_packed_weight_cache_invalidation(fake):
    """writes through .data bypass the version counter: invalidate_packed() is the documented way to repack"""
    m, _ = helpers.build_pair("pt")
    lr = m.cross_stage1
    pk0 = lr.packed()
    assert lr.packed() is pk0
    p = next(lr.parameters())
    p.data.mul_(2.0)                       # no version bump -> the cache cannot see it
    assert lr.packed() is pk0
    m.invalidate_packed()
    assert lr.packed() is not pk0
    pk1 = lr.packed()
    with torch.no_grad():
        p.mul_(0.5)                        # in-place op through the tensor: version bump -> repack
    assert lr.packed() is not pk1


def test_compat_registers_under_a_distinct_name_unless_overridden(monkeypatch):
    import sys
    import types
    from pcreid_b200 import compat, models

    class Reg(dict):
        def register_module(self, name, force, module):
            self[name] = module

    fake_builder = types.ModuleType("mmdet3d.models.builder")
    fake_builder.FUSIONMODELS = Reg(ReIDNet="the reference class")
    for n in ("mmdet3d", "mmdet3d.models"):
        monkeypatch.setitem(sys.modules, n, types.ModuleType(n))
    monkeypatch.setitem(sys.modules, "mmdet3d.models.builder", fake_builder)
    assert compat.install() is False
    assert fake_builder.FUSIONMODELS["ReIDNet"] == "the reference class" and fake_builder.FUSIONMODELS["ReIDNetB200"] is models.ReIDNet
    assert compat.install(override=True) is False
    assert fake_builder.FUSIONMODELS["ReIDNet"] is models.ReIDNet


def test_registry_and_state_dict_roundtrip():
    from pcreid_b200.models import FUSIONMODELS, build_model
    assert "ReIDNet" in FUSIONMODELS
    cfg = helpers.model_cfg("pt")
    m = build_model(cfg)
    assert cfg["backbone"]["type"] == "Pointnet_Backbone"          # cfg not mutated (the reference deletes 'type')
    m2 = build_model(helpers.model_cfg("pt"))
    m2.load_state_dict(m.state_dict(), strict=True)
    assert any(k.startswith("backbone.FP_modules.0.mlp_convs") for k in m.state_dict())   # dead reference weights kept


def test_train_mode_is_rejected(fake):
    m, _ = helpers.build_pair("pt")
    m.train()
    with pytest.raises(RuntimeError):
        m._encode(O.synth_objects(1, 128, 0))


def test_kernels_refuse_cpu_tensors():
    import pcreid_b200.kernels as K
    with pytest.raises(RuntimeError):
        K.cn_linear(torch.zeros(1, 4, 8), torch.zeros(4, 4))
    from pcreid_b200.ops import knn
    with pytest.raises(RuntimeError):
        knn(3, torch.zeros(1, 8, 3))


def test_feature_bank_and_reidentifier_match_the_tracker_semantics(fake):
    """PointFeatureSet.store_new / replace_old (tracking_feature_set.py:36-63) and the class-gated all-pairs scoring
    (tracking_point_reid.py:15-33, 95-116) against the oracle's pair-list formulation."""
    from pcreid_b200.models.tracking import PointFeatureSet, PointReidentifier, class_gate
    m, orc = helpers.build_pair("pt")
    bank = PointFeatureSet(replace_all=False)
    t0 = O.synth_objects(4, 128, 0)
    xyz, h = m.encode(t0)
    bank.store_new(xyz, h, torch.tensor([50, 10, 3, 1]))
    assert len(bank) == 4
    # a denser observation replaces track 1, a sparser one does not replace track 0
    t1 = O.synth_objects(2, 128, 5)
    x1, h1 = m.encode(t1)
    old0 = bank.pts_feats[0].clone()
    bank.replace_old(torch.tensor([0, 1]), x1, h1, torch.tensor([20, 30]))
    assert torch.equal(bank.pts_feats[0], old0) and torch.equal(bank.pts_feats[1], h1[1]) and bank.lengths.tolist() == [50, 30, 3, 1]
    dets = O.synth_objects(3, 128, 7)
    det_labels, det_len = torch.tensor([1, 2, 1]), torch.tensor([9, 9, 1])
    track_labels = torch.tensor([1, 1, 2, 1])
    reid = PointReidentifier(m, bank)
    cost, xd, hd = reid(dets, det_labels, det_len, torch.arange(4), track_labels)
    mask = class_gate(det_labels, track_labels, det_len, bank.lengths)
    assert mask.tolist() == O.class_gated_pairs(track_labels, bank.lengths, det_labels, det_len).tolist()
    assert mask.tolist() == [[True, False, False], [True, False, False], [False, True, False], [False, False, False]]
    xt, ht = bank.get_features(torch.arange(4))
    ref = orc.match_all_pairs(ht, xt, hd, xd, pair_mask=mask)
    assert (cost - ref).abs().max() < 2e-5 and (cost[~mask] == 0).all()


def test_pointnet_module_family_keys_and_builder():
    """state_dict keys follow mmcv's ConvModule naming so that mmdet3d checkpoints of these modules load unchanged
    (point_sa_module.py:283-300, point_fp_module.py:24-37); builder errors as pointnet_modules/builder.py:6-38."""
    from pcreid_b200.ops import PointFPModule, PointSAModule, PointSAModuleMSG, build_sa_module
    m = PointSAModuleMSG(num_point=16, radii=[0.5, 1.0], sample_nums=[8, 16], mlp_channels=[[4, 16, 32], [4, 32, 64]])
    keys = list(m.state_dict().keys())
    assert keys[:6] == ['mlps.0.layer0.conv.weight', 'mlps.0.layer0.bn.weight', 'mlps.0.layer0.bn.bias',
                        'mlps.0.layer0.bn.running_mean', 'mlps.0.layer0.bn.running_var', 'mlps.0.layer0.bn.num_batches_tracked']
    assert m.mlps[0].layer0.conv.weight.shape == (16, 7, 1, 1) and m.mlps[0].layer0.conv.bias is None
    f = PointFPModule([96, 64, 32])
    assert list(f.state_dict().keys())[0] == 'mlps.layer0.conv.weight'
    assert isinstance(build_sa_module(None, mlp_channels=[3, 8]), PointSAModule)
    with pytest.raises(KeyError):
        build_sa_module(dict(type="PAConvSAModule"))
    with pytest.raises(TypeError):
        build_sa_module([1, 2])
    nobn = PointSAModule(mlp_channels=[3, 8], num_point=4, radius=1.0, num_sample=4, norm_cfg=None)
    assert nobn.mlps[0].layer0.conv.bias is not None and not hasattr(nobn.mlps[0].layer0, "bn")


def test_pointnet_module_oracle_is_self_consistent():
    """CPU restatement: GroupAll pooling equals a plain max over per-point MLP outputs; FP module with a single source
    point copies that point's features (weights 1, 0, 0 -> the reference kernel's first-three-slots rule)."""
    from oracle import pointnet_modules_oracle as PO
    from pcreid_b200.ops import PointSAModule
    torch.manual_seed(0)
    g = PointSAModule(mlp_channels=[5, 8, 16]).eval()
    sd = g.state_dict()
    xyz, feat = O.synth_objects(2, 30, 1), torch.randn(2, 5, 30)
    _, gf, _ = PO.sa_module_msg(sd, None, [None], [None], xyz, feat)
    x = torch.cat([xyz.transpose(1, 2), feat], 1).unsqueeze(2)
    ref = PO._mlp(sd, "mlps.0", x).max(-1)[0]
    assert torch.equal(gf, ref)


# ---------------------------------------------------------------------------------------------------------------
# image-token matcher: cross_lin_attn + the token side of ImageReIDNet
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("S", [198, 64])
def test_image_token_matcher_vs_oracle_with_spec_kernels(fake, S):
    m, orc = helpers.build_image_pair()
    raw_t, raw_d = O.synth_tokens(3, 192, S, 0), O.synth_tokens(4, 192, S, 1)
    h_t, h_d = m.downsample_tokens(raw_t), m.downsample_tokens(raw_d)
    assert h_t.shape == (3, 64, S)
    assert (h_t - orc.downsample_tokens(raw_t)).abs().max() < 2e-5       # incl. the reference's un-permuted reshape
    assert (h_d - orc.downsample_tokens(raw_d)).abs().max() < 2e-5
    o_t, o_d = orc.downsample_tokens(raw_t), orc.downsample_tokens(raw_d)
    lg = m.match_forward_inference(h_t, h_d[:3])
    assert (lg - orc.match_forward_inference(o_t, o_d[:3])).abs().max() < 2e-5
    out = m.xcorr_eff(h_t, h_d[:3])
    assert out.shape == (3, 64, 2 * S) and (out - O.image_xcorr_eff(orc.sd, o_t, o_d[:3])).abs().max() < 2e-5
    assert (m.get_pooled_feats(out) - orc.pooled(O.image_xcorr_eff(orc.sd, o_t, o_d[:3]))).abs().max() < 2e-5
    L = m.match_all_pairs(h_t, h_d, chunk=5)
    Lo = orc.match_all_pairs(o_t, o_d)
    assert (L - Lo).abs().max() < 2e-5
    mask = torch.rand(3, 4, generator=torch.Generator().manual_seed(0)) > 0.4
    Lm = m.match_all_pairs(h_t, h_d, pair_mask=mask, chunk=5)
    assert (Lm - orc.match_all_pairs(o_t, o_d, pair_mask=mask)).abs().max() < 2e-5 and (Lm[~mask] == 0).all()


def test_image_reid_api_surface(fake):
    """registry name, state_dict keys of the reference (unused pos_mlp of cross_lin_attn included), forward_test with an
    attached token backbone, loud failure without one."""
    from pcreid_b200.models import FUSIONMODELS, ImageReIDNet
    assert "ImageReIDNet" in FUSIONMODELS
    m, orc = helpers.build_image_pair()
    keys = set(m.state_dict())
    assert {"cross_stage1.pos_mlp.0.weight", "cross_stage2.merge.weight", "downsample.2.bias", "vis_head.1.weight",
            "match_head.0.norm1.weight"} <= keys
    imgs = torch.randn(3, 3, 8, 8)
    with pytest.raises(RuntimeError):
        m.siamese_forward(imgs, imgs)

    class Out:
        def __init__(self, t):
            self.hidden_states = (None, t)

    class ToyBackbone(torch.nn.Module):            # stands in for DeiT: (B, 3, 8, 8) -> 198 tokens x 192
        def __init__(self):
            super().__init__()
            self.proj = torch.nn.Linear(192, 198 * 192)

        def forward(self, pixel_values):
            return Out(self.proj(pixel_values.reshape(pixel_values.shape[0], -1)).reshape(-1, 198, 192))

    torch.manual_seed(1)
    m.set_backbone(ToyBackbone(), name="deit-toy")
    s1, s2 = torch.randn(3, 3, 8, 8), torch.randn(3, 3, 8, 8)
    h1, h2 = m.siamese_forward(s1, s2)
    assert h1.shape == (3, 192, 198)
    one = lambda v: [torch.tensor([x]) for x in v]
    res = m(return_loss=False, sparse_1=list(s1), sparse_2=list(s2), label_1=one([1, 2, 12]), label_2=one([1, 2, 3]),
            vis_1=one([0, 1, -1]), vis_2=one([2, 3, 1]), id_1=one([5, 6, 7]), id_2=one([5, 9, 7]), size_1=one([4, 4, 4]),
            size_2=one([4, 4, 4]))[0]
    h_cat = torch.cat([h1, h2], 0)
    temp = orc.downsample_tokens(h_cat)
    assert (res['val_match_preds'] - orc.match_forward_inference(temp[:3], temp[3:])).abs().max() < 2e-5
    assert res['val_match_gt'].tolist() == [1., 0., 1.]
    assert res['val_cls_preds'].shape == (6, 20) and res['val_fp_preds'].shape == (6,) and res['val_vis_preds'].shape == (5, 4)
    sd = orc.sd
    cls_ref = O._lin(sd, "cls_head.1", O.linear_res(sd, "cls_head.0", orc.pooled(h_cat), 64))
    assert (res['val_cls_preds'] - cls_ref).abs().max() < 2e-5
    assert res['val_fp_gt'].tolist() == [0., 0., 1., 0., 0., 0.]
    with pytest.raises(NotImplementedError):
        m(return_loss=True)


def test_product_synthetic_module_matches_the_oracle_generators():
    """bench.py's measured legs draw inputs / configs from pcreid_b200.synthetic (never from oracle/ or tests/)"""
    from pcreid_b200 import synthetic as S
    for dup in (False, True):
        assert torch.equal(S.synth_objects(3, 64, 5, dup=dup), O.synth_objects(3, 64, 5, dup=dup))
    assert torch.equal(S.synth_tokens(2, 8, 9, 1), O.synth_tokens(2, 8, 9, 1))
    assert S.point_transformer_cfg((256, 128, 64)) == helpers.model_cfg("pt", (256, 128, 64))


def test_sa_layer_ball_query_grouping_vs_oracle_with_spec_kernels(fake):
    """PointNetSetAbstractionEdgeSA(use_knn=False) groups with query_ball_point (pointnet2_utils.py:218-240)"""
    from pcreid_b200.models.pointnet2_utils import PointNetSetAbstractionEdgeSA
    torch.manual_seed(66)
    sa = PointNetSetAbstractionEdgeSA(npoint=None, radius=1.2, nsample=24, mlp=[0, 32, 32, 32], sampling="RANDOM", use_xyz=True,
                                      use_knn=False).eval()
    sd = O.perturb_norm_state({"sa." + k: v for k, v in sa.state_dict().items()})
    sa.load_state_dict({k[3:]: v for k, v in sd.items()})
    x = O.synth_objects(2, 128, 3)
    with torch.no_grad():
        nx, nf = sa(x, None, 64)
    ox, of = O.sa_layer(sd, "sa", x, None, 64, 24, radius=1.2)
    assert torch.equal(nx, ox) and (nf - of).abs().max() < 2e-5
    import fake_kernels
    q = x[:, :64].contiguous()
    assert torch.equal(fake_kernels.query_ball_point(0.7, 16, x, q).long(), O.query_ball_point(0.7, 16, x, q))


@pytest.mark.parametrize("use_knn", [True, False])
def test_sa_layer_fps_sampling_vs_oracle_with_spec_kernels(fake, use_knn):
    """PointNetSetAbstractionEdgeSA(sampling='FPS'): centres gathered by the sampled indices"""
    from pcreid_b200.models.pointnet2_utils import PointNetSetAbstractionEdgeSA
    torch.manual_seed(66)
    sa = PointNetSetAbstractionEdgeSA(npoint=None, radius=1.5, nsample=24, mlp=[64, 64, 64, 64], sampling="FPS", use_xyz=True,
                                      use_knn=use_knn).eval()
    sd = O.perturb_norm_state({"sa." + k: v for k, v in sa.state_dict().items()})
    sa.load_state_dict({k[3:]: v for k, v in sd.items()})
    x, f = O.synth_objects(2, 128, 3), torch.randn(2, 32, 128)
    torch.manual_seed(5)
    with torch.no_grad():
        nx, nf = sa(x, f, 64)                       # draws the start indices from the host RNG like the reference
    torch.manual_seed(5)
    start = torch.randint(0, 128, (2,), dtype=torch.long)
    ox, of = O.sa_layer(sd, "sa", x, f, 64, 24, radius=None if use_knn else 1.5, fps_start=start)
    assert torch.equal(nx, ox) and (nf - of).abs().max() < 2e-5
