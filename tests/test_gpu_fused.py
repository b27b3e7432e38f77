"""-m gpu: fused tcgen05 matcher against the fp32 parity path and the oracle, in both tensor-core modes.
Tolerances (fp32 accumulate / norms; SURVEY.md 7 pre-study and 8d parity gates):
  'fast'      bf16 operands (8-bit significand):                 |dlogit| <= 3e-2
  'parity_tc' fp16 operands (11-bit significand, same as tf32):  |dlogit| <= 5e-3  (the tf32 gate)
and the top-1 decision must be unchanged on rows whose oracle gap exceeds 2x the measured error."""
import pytest
import torch

import helpers
from oracle import reid_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL_FAST = 3e-2
TOL_TC = 5e-3
TOL = {"fast": TOL_FAST, "parity_tc": TOL_TC}


@pytest.mark.parametrize("mode", ["fast", "parity_tc"])
@pytest.mark.parametrize("N,T,D", [(256, 5, 7), (128, 6, 4), (512, 2, 3), (160, 5, 4), (192, 3, 5), (224, 4, 4)])
def test_fused_matches_parity_and_oracle(N, T, D, mode):
    """point counts of the reference's ablation configs, including the ragged ones (160 / 192 / 224: zero-padded tiles)"""
    from pcreid_b200.models import fused_pairs
    m, orc = helpers.build_pair("pt", (N, N // 2, N // 4), device=DEV)
    t, d = O.synth_objects(T, N, 0), O.synth_objects(D, N, 1)
    xt, ht = m.encode(t.to(DEV))
    xd, hd = m.encode(d.to(DEV))
    Lp = m.match_all_pairs(ht, xt, hd, xd).cpu()
    m.match_mode = mode
    assert fused_pairs.supported(m, N)
    Lf = m.match_all_pairs(ht, xt, hd, xd).cpu()
    Lo = orc.match_all_pairs(ht.cpu(), xt.cpu(), hd.cpu(), xd.cpu())
    err = (Lf - Lo).abs().max().item()
    assert (Lp - Lo).abs().max() < 1e-4
    assert err < TOL[mode], f"{mode}-mode logits off by {err}"
    ok, agree, n = helpers.margin_aware_top1(Lo, Lf, err)
    assert ok, f"top-1 changed on a decisive row (agreement {agree}, {n} decisive rows)"


@pytest.mark.parametrize("mode", ["fast", "parity_tc"])
def test_fused_with_pair_mask_and_unsorted_pairs(mode):
    m, orc = helpers.build_pair("pt", (256, 128, 64), device=DEV)
    t, d = O.synth_objects(9, 256, 2), O.synth_objects(11, 256, 3)
    xt, ht = m.encode(t.to(DEV))
    xd, hd = m.encode(d.to(DEV))
    mask = torch.rand(9, 11, generator=torch.Generator().manual_seed(0)) > 0.5
    m.match_mode = mode
    Lf = m.match_all_pairs(ht, xt, hd, xd, pair_mask=mask.to(DEV)).cpu()
    Lo = orc.match_all_pairs(ht.cpu(), xt.cpu(), hd.cpu(), xd.cpu(), pair_mask=mask)
    assert (Lf - Lo).abs().max() < TOL[mode]
    assert (Lf[~mask] == 0).all()


@pytest.mark.parametrize("mode", ["fast", "parity_tc"])
def test_fused_symmetry_and_determinism(mode):
    m, _ = helpers.build_pair("pt", (256, 128, 64), device=DEV)
    a = O.synth_objects(8, 256, 4).to(DEV)
    xa, ha = m.encode(a)
    m.match_mode = mode
    L1 = m.match_all_pairs(ha, xa, ha, xa)
    L2 = m.match_all_pairs(ha, xa, ha, xa)
    assert torch.equal(L1, L2)
    assert (L1 - L1.t()).abs().max() < TOL[mode]


@pytest.mark.parametrize("mode", ["fast", "parity_tc"])
@pytest.mark.parametrize("N", [256, 128])
def test_full_tensor_core_mode_end_to_end(N, mode):
    """set_mode('fast' | 'parity_tc'): tf32 tensor-core SA MLPs / attention blocks in the encoder + fused bf16 / fp16 matcher,
    against the oracle end to end."""
    m, orc = helpers.build_pair("pt", (N, N // 2, N // 4), device=DEV)
    m.set_mode(mode)
    t, d = O.synth_objects(6, N, 10), O.synth_objects(5, N, 11)
    xt, ht = m.encode(t.to(DEV))
    xd, hd = m.encode(d.to(DEV))
    oxt, oht = orc.encode(t)
    oxd, ohd = orc.encode(d)
    assert (ht.cpu() - oht).abs().max() < 2e-2       # tf32 operands in the three SA shared MLPs
    Lf = m.match_all_pairs(ht, xt, hd, xd).cpu()
    Lo = orc.match_all_pairs(oht, oxt, ohd, oxd)
    err = (Lf - Lo).abs().max().item()
    assert err < TOL[mode], f"{mode}-mode logits off by {err}"
    ok, agree, n = helpers.margin_aware_top1(Lo, Lf, err)
    assert ok, f"top-1 changed on a decisive row (agreement {agree}, {n} decisive rows)"


@pytest.mark.parametrize("T,D,masked", [(37, 70, False), (5, 3, True), (130, 257, True)])
def test_concat_head_tensor_core_matches_parity_and_oracle(T, D, masked):
    """'concat' baseline head (reid_pts_point-transformer_baseline.py) on the tensor cores (csrc/concat_tc.cu): ragged T / D,
    class-gate mask, against the fp32 kernel and the oracle."""
    m, orc = helpers.build_pair("concat", (128, 64, 32), device=DEV)
    t, d = O.synth_objects(T, 128, 30), O.synth_objects(D, 128, 31)
    xt, ht = m.encode(t.to(DEV))
    xd, hd = m.encode(d.to(DEV))
    mask = (torch.rand(T, D, generator=torch.Generator().manual_seed(2)) > 0.4) if masked else None
    mk = None if mask is None else mask.to(DEV)
    Lp = m.match_all_pairs(ht, xt, hd, xd, pair_mask=mk).cpu()
    m.match_mode = 'fast'
    Lf = m.match_all_pairs(ht, xt, hd, xd, pair_mask=mk).cpu()
    Lo = orc.match_all_pairs(ht.cpu(), xt.cpu(), hd.cpu(), xd.cpu(), pair_mask=mask)
    assert (Lp - Lo).abs().max() < 1e-4
    err = (Lf - Lo).abs().max().item()
    assert err < TOL_FAST, f"tensor-core concat head off by {err}"
    if masked:
        assert (Lf[~mask] == 0).all()
    if T > 8:
        ok, agree, n = helpers.margin_aware_top1(Lo, Lf, err)
        assert ok, f"top-1 changed on a decisive row (agreement {agree}, {n} decisive rows)"


def test_full_size_config_properties():
    """BASELINE configs[1] at its full size (1024 tracks x 1024 detections x 256 points, fast mode): size-independent
    properties instead of an oracle pass -- (1) the dense row-chunked driver and the pair-list driver give the same logit for
    the same pair, (2) scoring a row block on its own equals the rows of the full matrix (what the row-sharded multi-GPU
    driver relies on), (3) swapping the roles of the two sets transposes the matrix (xcorr_eff + 'both' pooling is symmetric),
    (4) a sampled 6 x 6 block agrees with the oracle within the fast-mode tolerance."""
    T = D = 1024
    m, orc = helpers.build_pair("pt", (256, 128, 64), device=DEV, perturb=False)
    m.set_mode('fast')
    t, d = O.synth_objects(T, 256, 0), O.synth_objects(D, 256, 1)
    xt, ht = m.encode(t.to(DEV))
    xd, hd = m.encode(d.to(DEV))
    L = m.match_all_pairs(ht, xt, hd, xd)
    assert L.shape == (T, D) and torch.isfinite(L).all()
    mask = torch.rand(T, D, generator=torch.Generator().manual_seed(1)) < 2e-3
    Lm = m.match_all_pairs(ht, xt, hd, xd, pair_mask=mask.to(DEV))
    assert int(mask.sum()) > 1000 and torch.equal(Lm[mask.to(DEV)], L[mask.to(DEV)]) and (Lm[~mask.to(DEV)] == 0).all()
    rows = slice(384, 512)
    assert torch.equal(m.match_all_pairs(ht[rows], xt[rows], hd, xd), L[rows])
    Lt = m.match_all_pairs(hd[:256], xd[:256], ht[:256], xt[:256])
    assert (Lt - L[:256, :256].t()).abs().max() < TOL_FAST
    ti, di = torch.tensor([0, 1, 511, 512, 1022, 1023]), torch.tensor([3, 64, 65, 700, 1000, 1023])
    oxt, oht = orc.encode(t[ti])
    oxd, ohd = orc.encode(d[di])
    Lo = orc.match_all_pairs(oht, oxt, ohd, oxd)
    assert (L[ti][:, di].cpu() - Lo).abs().max() < TOL_FAST


@pytest.mark.parametrize("mode", ["fast", "parity_tc"])
def test_small_dense_match_replays_from_a_cuda_graph(mode):
    """enable_cuda_graphs(): a per-frame sized dense matrix (T*D <= GRAPH_MAX_PAIRS, no mask) is replayed from a captured graph --
    bit-identical to the eager launches, correct for new inputs of the same shape, masked / large calls stay eager"""
    m, _ = helpers.build_pair("pt", (128, 64, 32), device=DEV)
    m.set_mode(mode)
    outs = []
    for seed in (0, 5):
        t, d = O.synth_objects(9, 128, seed).to(DEV), O.synth_objects(12, 128, seed + 1).to(DEV)
        m.enable_cuda_graphs(False)
        xt, ht = m.encode(t)
        xd, hd = m.encode(d)
        eager = m.match_all_pairs(ht, xt, hd, xd)
        m.enable_cuda_graphs(True)
        xt2, ht2 = m.encode(t)
        xd2, hd2 = m.encode(d)
        assert torch.equal(ht, ht2) and torch.equal(hd, hd2)
        got = m.match_all_pairs(ht2, xt2, hd2, xd2)
        assert torch.equal(got, eager)
        outs.append(got)
        mask = torch.rand(9, 12, generator=torch.Generator().manual_seed(3)) > 0.4
        assert torch.equal(m.match_all_pairs(ht2, xt2, hd2, xd2, pair_mask=mask.to(DEV)), eager * mask.to(DEV))
    assert any(k[0] == "match" for k in m._graphs if isinstance(k[0], str))
    assert not torch.equal(outs[0], outs[1])


_AB_SCRIPT = r'''
import sys, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import helpers
from oracle import reid_oracle as O
out = {}
for mode in ("parity_tc", "fast"):
    for N, T, D in ((256, 9, 7), (160, 4, 5), (128, 6, 3)):
        m, _ = helpers.build_pair("pt", (N, N // 2, N // 4), device="cuda")
        m.set_mode(mode)
        xt, ht = m.encode(O.synth_objects(T, N, 0).cuda())
        xd, hd = m.encode(O.synth_objects(D, N, 1).cuda())
        out[f"{mode}_{N}"] = m.match_all_pairs(ht, xt, hd, xd).cpu()
        mask = torch.rand(T, D, generator=torch.Generator().manual_seed(N)) > 0.4
        out[f"{mode}_{N}_masked"] = m.match_all_pairs(ht, xt, hd, xd, pair_mask=mask.cuda()).cpu()
torch.save(out, sys.argv[2])
'''


def test_ab_kernel_variants_are_bit_identical(tmp_path):
    """The kernels this library shipped before the shared-memory-pipe work stay selectable for A/B runs (PCREID_X_TMEM=0: X' and the
    bias operand as shared-memory images; PCREID_P1B=1: phase 1b with `a` as a cp.async image and an M = 128 key/value GEMM;
    PCREID_P1B_ORDER=templ: one unit order for both phase-1 kernels).  Every variant computes the same sums in the same order: the
    logits must be bit-identical to the default build's, dense and pair-list drivers, full and ragged tiles, both operand formats."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "ab.py"
    script.write_text(_AB_SCRIPT)
    res = {}
    for tag, env in (("default", {}), ("before", {"PCREID_X_TMEM": "0", "PCREID_P1B": "1", "PCREID_P1B_ORDER": "templ"}),
                     ("split", {"PCREID_P1B": "1", "PCREID_P1B_SPLIT": "2"})):
        out = tmp_path / f"{tag}.pt"
        e = dict(os.environ)
        e.update(env)
        subprocess.run([sys.executable, str(script), root, str(out)], check=True, env=e, timeout=600)
        res[tag] = torch.load(out)
    for tag in ("before", "split"):
        for k, v in res["default"].items():
            assert torch.equal(v, res[tag][k]), f"{tag} variant differs from the default kernels on {k}"


@pytest.mark.parametrize("mode", ["parity_tc", "fast"])
@pytest.mark.parametrize("N,T,D", [(1024, 3, 4), (640, 3, 3), (96, 5, 6)])
def test_fused_long_and_short_objects(N, T, D, mode):
    """tile counts beyond the headline's two: 8 tiles per object (the 1024-point Waymo config), 5 tiles, and a single ragged tile --
    the row prefetch of phase 1b runs two tiles ahead across unit boundaries, the key/value GEMM accumulates over all of them;
    dense driver and pair-list driver must agree bit for bit on the pairs they share"""
    m, _ = helpers.build_pair("pt", (N, N // 2, N // 4), device=DEV)
    xt, ht = m.encode(O.synth_objects(T, N, 0).to(DEV))
    xd, hd = m.encode(O.synth_objects(D, N, 1).to(DEV))
    Lp = m.match_all_pairs(ht, xt, hd, xd).cpu()
    m.match_mode = mode
    Lf = m.match_all_pairs(ht, xt, hd, xd).cpu()
    assert (Lf - Lp).abs().max() < TOL[mode]
    mask = torch.rand(T, D, generator=torch.Generator().manual_seed(1)) > 0.3
    Lm = m.match_all_pairs(ht, xt, hd, xd, pair_mask=mask.to(DEV)).cpu()
    assert torch.equal(Lm[mask], Lf[mask]) and float(Lm[~mask].abs().max()) == 0.0
