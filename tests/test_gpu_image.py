"""-m gpu: image-token matcher -- cross_lin_attn + the token side of ImageReIDNet (downsample, xcorr_eff,
pooling, match head, all-pairs driver) on the C-ABI kernels, against the oracle and the committed golden vectors generated
from the reference modules.  fp32 parity mode: 1e-4; fast mode (fused bf16 tcgen05 matcher with a zero position image):
|dlogit| <= 3e-2 and the margin-aware top-1 check."""
import pytest
import torch

import helpers
from oracle import reid_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-4
TOL_FAST = 3e-2


def test_cross_lin_attn_module_vs_oracle():
    from pcreid_b200.models import cross_lin_attn
    torch.manual_seed(66)
    ca = cross_lin_attn(64, 2).eval()
    sd = O.perturb_norm_state({"x." + k: v for k, v in ca.state_dict().items()})
    ca.load_state_dict({k[2:]: v for k, v in sd.items()})
    ca = ca.to(DEV)
    a, b = O.synth_tokens(3, 64, 198, 0), O.synth_tokens(3, 64, 77, 1)       # ragged, different lengths on the two sides
    got = ca(a.to(DEV), b.to(DEV)).cpu()
    assert (got - O.cross_lin_attention(sd, "x", a, b)).abs().max() < TOL


@pytest.mark.parametrize("S", [198, 128])
def test_image_tokens_parity_vs_oracle(S):
    m, orc = helpers.build_image_pair(device=DEV)
    raw = O.synth_tokens(5, 192, S, 0)
    h = m.downsample_tokens(raw.to(DEV))
    ho = orc.downsample_tokens(raw)
    assert (h.cpu() - ho).abs().max() < TOL
    L = m.match_all_pairs(h[:2], h[2:]).cpu()
    assert (L - orc.match_all_pairs(ho[:2], ho[2:])).abs().max() < TOL
    lg = m.match_forward_inference(h[:2], h[2:4]).cpu()
    assert (lg - orc.match_forward_inference(ho[:2], ho[2:4])).abs().max() < TOL


def test_image_tokens_vs_reference_golden():
    g = helpers.golden("reid_image_tokens")
    m, _ = helpers.build_image_pair(device=DEV)
    assert abs(helpers.weight_checksum({k: v.cpu() for k, v in m.state_dict().items()}) - float(g["weight_checksum"])) \
        < 1e-6 * float(g["weight_checksum"])
    h = m.downsample_tokens(torch.from_numpy(g["raw"]).to(DEV)).cpu()
    assert (h - torch.from_numpy(g["h_raw"])).abs().max() < TOL
    L = m.match_all_pairs(torch.from_numpy(g["h_t"]).to(DEV), torch.from_numpy(g["h_d"]).to(DEV)).cpu()
    assert (L - torch.from_numpy(g["logits"])).abs().max() < TOL
    m.set_mode('fast')
    Lf = m.match_all_pairs(torch.from_numpy(g["h_t"]).to(DEV), torch.from_numpy(g["h_d"]).to(DEV)).cpu()
    assert (Lf - torch.from_numpy(g["logits"])).abs().max() < TOL_FAST


@pytest.mark.parametrize("S,T,D", [(198, 6, 5), (256, 4, 4), (64, 3, 7)])
def test_image_tokens_fast_mode(S, T, D):
    """fused tcgen05 matcher on token sets: 198 tokens (DeiT-distilled @224) = one full + one 70-row tile"""
    from pcreid_b200.models import fused_pairs
    m, orc = helpers.build_image_pair(device=DEV)
    assert fused_pairs.supported(m, S)
    h_t, h_d = O.synth_tokens(T, 64, S, 3), O.synth_tokens(D, 64, S, 4)
    Lo = orc.match_all_pairs(h_t, h_d)
    m.set_mode('fast')
    Lf = m.match_all_pairs(h_t.to(DEV), h_d.to(DEV)).cpu()
    err = (Lf - Lo).abs().max().item()
    assert err < TOL_FAST, f"fast-mode logits off by {err}"
    ok, agree, n = helpers.margin_aware_top1(Lo, Lf, err)
    assert ok, f"top-1 changed on a decisive row (agreement {agree}, {n} decisive rows)"
    mask = torch.rand(T, D, generator=torch.Generator().manual_seed(0)) > 0.5
    Lm = m.match_all_pairs(h_t.to(DEV), h_d.to(DEV), pair_mask=mask.to(DEV)).cpu()
    assert (Lm - orc.match_all_pairs(h_t, h_d, pair_mask=mask)).abs().max() < TOL_FAST and (Lm[~mask] == 0).all()
