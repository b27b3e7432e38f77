"""-m gpu: the encoders and match heads through the reference-facing model API on cuda:0, against the oracle on
the same seeded inputs and against the committed golden vectors (tests/golden, made from the reference modules).

Tolerances (fp32 parity mode): features 1e-4 absolute (values are O(1)), logits 1e-4, top-1 ReID
decision unchanged on every row whose oracle top-1/top-2 gap exceeds 2x the tolerance."""
import pytest
import torch

import helpers
import pcreid_b200.kernels as K
from oracle import reid_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-4


@pytest.mark.parametrize("name,kind", [("reid_pt", "pt"), ("reid_pt256", "pt"), ("reid_dgcnn", "dgcnn"), ("reid_pointnet", "pointnet"),
                                       ("reid_xcorr", "xcorr"), ("reid_xcorr-baseline", "xcorr-baseline")])
def test_golden_vectors(name, kind):
    g = helpers.golden(name)
    m, _ = helpers.build_pair(kind, tuple(int(v) for v in g["backbone_list"]), device=DEV)
    assert abs(helpers.weight_checksum(m.state_dict()) - float(g["weight_checksum"])) < 1e-6 * float(g["weight_checksum"])
    t, d = torch.from_numpy(g["tracks"]).to(DEV), torch.from_numpy(g["dets"]).to(DEV)
    with torch.no_grad():
        xt, ht = m._encode(t)
        xd, hd = m._encode(d)
    assert (ht.cpu() - torch.from_numpy(g["h_t"])).abs().max() < TOL
    assert (hd.cpu() - torch.from_numpy(g["h_d"])).abs().max() < TOL
    L = m.match_all_pairs(ht, xt, hd, xd).cpu()
    ref = torch.from_numpy(g["logits"])
    assert (L - ref).abs().max() < TOL
    ok, agree, n = helpers.margin_aware_top1(ref, L, TOL)
    assert ok, f"top-1 changed on a decisive row (raw agreement {agree}, {n} decisive rows)"


@pytest.mark.parametrize("kind,N,blist", [("pt", 128, (128, 64, 32)), ("pt", 256, (256, 128, 64)), ("pt", 160, (160, 80, 40)),
                                          ("concat", 128, (128, 64, 32)), ("dgcnn", 256, (128, 64, 32)),
                                          ("pointnet", 128, (128, 64, 32)),
                                          ("pt15m", 128, (128, 64, 32)), ("pt7m", 128, (128, 64, 32)),
                                          ("xcorr", 128, (128, 64, 32)), ("xcorr-baseline", 256, (256, 128, 64))])
@pytest.mark.parametrize("dup", [False, True])
def test_model_vs_oracle(kind, N, blist, dup):
    m, orc = helpers.build_pair(kind, blist, device=DEV)
    t, d = O.synth_objects(6, N, 0, dup=dup), O.synth_objects(7, N, 1, dup=dup)
    with torch.no_grad():
        xt, ht = m._encode(t.to(DEV))
        xd, hd = m._encode(d.to(DEV))
    oxt, oht = orc.encode(t)
    oxd, ohd = orc.encode(d)
    assert ht.shape == oht.shape and ht.is_contiguous()
    assert (ht.cpu() - oht).abs().max() < TOL and (hd.cpu() - ohd).abs().max() < TOL
    L = m.match_all_pairs(ht, xt, hd, xd, chunk=16).cpu()
    Lo = orc.match_all_pairs(oht, oxt, ohd, oxd)
    assert (L - Lo).abs().max() < TOL
    ok, agree, n = helpers.margin_aware_top1(Lo, L, TOL)
    assert ok, f"top-1 changed on a decisive row (raw agreement {agree}, {n} decisive rows)"
    mask = torch.rand(6, 7, generator=torch.Generator().manual_seed(5)) > 0.5
    Lm = m.match_all_pairs(ht, xt, hd, xd, pair_mask=mask.to(DEV)).cpu()
    assert (Lm - Lo * mask).abs().max() < TOL


@pytest.mark.parametrize("kind,N,blist", [("dgcnn", 256, (128, 64, 32)), ("pointnet", 128, (128, 64, 32)), ("pt15m", 128, (128, 64, 32)),
                                          ("pt7m", 128, (128, 64, 32)), ("xcorr", 128, (128, 64, 32)),
                                          ("xcorr-baseline", 256, (256, 128, 64))])
@pytest.mark.parametrize("mode,tol", [("fast", 3e-2), ("parity_tc", 5e-3)])
def test_tensor_core_modes_every_config(kind, N, blist, mode, tol):
    """the tensor-core modes on the configs the fused xcorr_eff matcher does not cover: DGCNN EdgeConv / conv5 / downsample and the
    PointNet shared MLPs as TMA-staged tcgen05 kind::tf32 GEMMs (cn_linear_tma.cu; every contraction in 'fast' mode, K >= 256 in
    'parity_tc'), the unfused cross-attention chains of the d_model = 128 / 'xcorr' / 'xcorr-baseline' matchers on the same kernel
    with the pair gather maps as tensor-map coordinates.  Gates: |dlogit| within the mode's tolerance, decisive rows keep their top-1."""
    m, orc = helpers.build_pair(kind, blist, device=DEV)
    m.set_mode(mode)
    t, d = O.synth_objects(6, N, 0), O.synth_objects(7, N, 1)
    used = {"tma": 0}
    real = K._OPS.cn_linear_tma

    class _Count:
        def __getattr__(self, name):
            if name == "cn_linear_tma":
                def f(*a):
                    rc = real(*a)
                    used["tma"] += rc == 0
                    return rc
                return f
            return getattr(ops, name)
    ops = K._OPS
    K._OPS = _Count()
    try:
        xt, ht = m.encode(t.to(DEV))
        xd, hd = m.encode(d.to(DEV))
        L = m.match_all_pairs(ht, xt, hd, xd).cpu()
    finally:
        K._OPS = ops
    assert used["tma"] > 0, "no GEMM of this configuration reached the TMA-staged tensor-core kernel"
    oxt, oht = orc.encode(t)
    oxd, ohd = orc.encode(d)
    Lo = orc.match_all_pairs(oht, oxt, ohd, oxd)
    scale = max(1.0, float(oht.abs().max()))
    assert float((ht.cpu() - oht).abs().max()) < 2e-2 * scale
    err = float((L - Lo).abs().max())
    assert err < tol, err
    ok, agree, n = helpers.margin_aware_top1(Lo, L, err)
    assert ok, f"top-1 changed on a decisive row (raw agreement {agree}, {n} decisive rows)"


@pytest.mark.parametrize("kind,N,blist", [("pt", 256, (256, 128, 64)), ("dgcnn", 256, (128, 64, 32)), ("pointnet", 128, (128, 64, 32)),
                                          ("pt7m", 128, (128, 64, 32)), ("xcorr", 128, (128, 64, 32))])
def test_parity_x3_mode_meets_the_fp32_gate(kind, N, blist):
    """set_mode('parity_x3'): the strict-parity path with its contractions on tcgen05 (3 x tf32) -- same 1e-4 gates as fp32 parity"""
    m, orc = helpers.build_pair(kind, blist, device=DEV)
    m.set_mode('parity_x3')
    t, d = O.synth_objects(6, N, 0), O.synth_objects(7, N, 1)
    xt, ht = m.encode(t.to(DEV))
    xd, hd = m.encode(d.to(DEV))
    oxt, oht = orc.encode(t)
    oxd, ohd = orc.encode(d)
    assert (ht.cpu() - oht).abs().max() < TOL and (hd.cpu() - ohd).abs().max() < TOL
    L = m.match_all_pairs(ht, xt, hd, xd, chunk=16).cpu()
    Lo = orc.match_all_pairs(oht, oxt, ohd, oxd)
    assert (L - Lo).abs().max() < TOL
    ok, agree, n = helpers.margin_aware_top1(Lo, L, TOL)
    assert ok, f"top-1 changed on a decisive row (raw agreement {agree}, {n} decisive rows)"


def test_degenerate_clouds_all_points_identical():
    """all-zero / single-point clouds (empty crops, pc_utils.py:84-89): every kNN distance ties; features must still
    match because tied neighbours are identical points."""
    m, orc = helpers.build_pair("pt", device=DEV)
    t = torch.zeros(2, 128, 3)
    t[1] = O.synth_objects(1, 1, 3)[0, 0]
    d = O.synth_objects(3, 128, 1, dup=True)
    with torch.no_grad():
        xt, ht = m._encode(t.to(DEV))
        xd, hd = m._encode(d.to(DEV))
    oxt, oht = orc.encode(t)
    oxd, ohd = orc.encode(d)
    assert (ht.cpu() - oht).abs().max() < TOL
    assert (m.match_all_pairs(ht, xt, hd, xd).cpu() - orc.match_all_pairs(oht, oxt, ohd, oxd)).abs().max() < TOL


def test_reference_api_on_gpu():
    m, orc = helpers.build_pair("pt", device=DEV)
    s1, s2 = O.synth_objects(4, 128, 2), O.synth_objects(4, 128, 3)
    x1, x2, h1, h2 = m.siamese_forward(s1.to(DEV), s2.to(DEV))
    o = orc.siamese_forward(s1, s2)
    lg = m.match_forward_inference(h1, h2, x1, x2)
    assert (lg.cpu() - orc.match_forward_inference(o[2], o[3], o[0], o[1])).abs().max() < TOL
    out, o1, o2 = m.xcorr_eff(h1, x1, h2, x2)
    ro, _, _ = O.xcorr_eff(orc.sd, o[2], o[0], o[3], o[1])
    assert (out.cpu() - ro).abs().max() < TOL
    assert (m.get_pooled_feats(out).cpu() - O.pooled_feats(ro, "both")).abs().max() < TOL
    xyz, feat = m.forward_inference(s1.to(DEV))
    assert (feat.cpu() - o[2]).abs().max() < TOL


def test_all_pairs_symmetry_property():
    """point-cat + max/avg pooling makes the logit symmetric under swapping the two objects."""
    m, _ = helpers.build_pair("pt", device=DEV)
    a = O.synth_objects(9, 128, 4).to(DEV)
    with torch.no_grad():
        xa, ha = m._encode(a)
    L = m.match_all_pairs(ha, xa, ha, xa)
    assert (L - L.t()).abs().max() < 1e-4


def test_missing_library_fails_loudly(monkeypatch):
    from pcreid_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libpcreid_sm100.so")
    import pcreid_b200.build as bld
    monkeypatch.setattr(bld, "build_library", lambda *a, **k: None)
    with pytest.raises((RuntimeError, OSError)):
        _lib.lib()


def test_cuda_graph_encode_is_identical_and_survives_weight_updates():
    """encode() through a captured CUDA graph: bit-identical to the eager launches, results stay valid across calls, the
    graph is rebuilt when a parameter changes"""
    m, orc = helpers.build_pair("pt", (128, 64, 32), device=DEV)
    m.set_mode('fast')
    a, b = O.synth_objects(9, 128, 40).to(DEV), O.synth_objects(9, 128, 41).to(DEV)
    xa, ha = m.encode(a)
    xb, hb = m.encode(b)
    m.enable_cuda_graphs(True)
    xa2, ha2 = m.encode(a)
    xb2, hb2 = m.encode(b)                      # same shape: replays the same graph, must not clobber (xa2, ha2)
    assert torch.equal(ha, ha2) and torch.equal(hb, hb2) and torch.equal(xa, xa2)
    with torch.no_grad():
        m.backbone.cov_final.bias.add_(0.5)
    _, ha3 = m.encode(a)
    assert (ha3 - (ha + 0.5)).abs().max() < 1e-5
    m.enable_cuda_graphs(False)
    assert torch.equal(m.encode(a)[1], ha3)
