"""-m gpu: fused tensor-core linear-attention blocks (csrc/attn_tc.cu: attn_front / kv_merge / attn_back) against the
fp32 kernel chain of the same module (itself pinned to the oracle by test_gpu_model) and against plain torch.
Tolerance: tf32 operands (10-bit mantissa), fp32 accumulation and LayerNorms -> 1e-2 absolute on LayerNorm-scaled
outputs (measured ~2e-3)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-2


def _perturb(mod, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in mod.named_parameters():
            if "norm" in n:
                p.add_(0.2 * torch.randn(p.shape, generator=g))
    return mod


def _torch_self_attention(m, feat, xyz):
    """plain torch restatement of Self_Attention.forward (mmdet3d/models/pointnet2_utils.py:86-114)."""
    bs = feat.shape[0]
    f = feat.permute(0, 2, 1)
    fp = f + m.pos_mlp(xyz)
    q = m.q_proj(fp).view(bs, -1, m.nhead, m.dim)
    k = m.k_proj(fp).view(bs, -1, m.nhead, m.dim)
    v = m.v_proj(fp).view(bs, -1, m.nhead, m.dim)
    Q, Kf = torch.nn.functional.elu(q) + 1, torch.nn.functional.elu(k) + 1
    L = v.shape[1]
    KV = torch.einsum("nshd,nshv->nhdv", Kf, v / L)
    Z = 1 / (torch.einsum("nlhd,nhd->nlh", Q, Kf.sum(1)) + 1e-6)
    msg = torch.einsum("nlhd,nhdv,nlh->nlhv", Q, KV, Z) * L
    msg = m.norm1(m.merge(msg.reshape(bs, -1, m.nhead * m.dim)))
    msg = m.norm2(m.mlp(torch.cat([f, msg], 2)))
    return (f + msg).permute(0, 2, 1)


@pytest.mark.parametrize("d,S,B", [(32, 256, 5), (64, 128, 4), (128, 64, 3), (32, 160, 3), (64, 80, 3), (128, 40, 2), (64, 300, 2)])
def test_self_attention_block(d, S, B):
    from pcreid_b200.models.pointnet2_utils import Self_Attention
    torch.manual_seed(d + S)
    m = _perturb(Self_Attention(d, 2), 1).to(DEV).eval()
    feat = torch.randn(B, d, S, device=DEV)
    xyz = torch.randn(B, S, 3, device=DEV) * torch.tensor([2.0, 0.9, 0.8], device=DEV)
    with torch.no_grad():
        ref = m(feat, xyz)
        gold = _torch_self_attention(m, feat, xyz)
        m.tc_mode = True
        got = m(feat, xyz)
    assert (ref - gold).abs().max() < 2e-4
    err = (got - gold).abs().max().item()
    assert err < TOL, f"fused Self_Attention d={d} S={S}: max err {err}"


@pytest.mark.parametrize("f1,f2,d,out,N,S,pm", [(64, 128, 64, 128, 128, 64, False), (32, 128, 64, 64, 256, 128, False),
                                               (3, 64, 64, 32, 256, 256, True), (64, 128, 64, 128, 80, 40, False),
                                               (3, 64, 64, 32, 160, 160, True)])
def test_fp_sa_block(f1, f2, d, out, N, S, pm):
    from pcreid_b200.models.pointnet2_utils import FP_SA
    torch.manual_seed(f1 + N)
    m = _perturb(FP_SA(0, f1, f2, d, out, 2), 2).to(DEV).eval()
    B = 3
    scale = torch.tensor([2.0, 0.9, 0.8], device=DEV)
    xyz1, xyz2 = torch.randn(B, N, 3, device=DEV) * scale, torch.randn(B, S, 3, device=DEV) * scale
    feat1 = xyz1.contiguous() if pm else torch.randn(B, f1, N, device=DEV)
    feat2 = torch.randn(B, f2, S, device=DEV)
    with torch.no_grad():
        ref = m(feat1, xyz1, feat2, xyz2, feat1_point_major=pm)
        m.tc_mode = True
        got = m(feat1, xyz1, feat2, xyz2, feat1_point_major=pm)
    assert got.shape == ref.shape == (B, out, N)
    err = (got - ref).abs().max().item()
    assert err < TOL, f"fused FP_SA {f1, f2, d, out, N, S}: max err {err}"


def test_fast_encoder_close_to_parity_encoder():
    import helpers
    from oracle import reid_oracle as O
    m, _ = helpers.build_pair("pt", (256, 128, 64), device=DEV)
    x = O.synth_objects(6, 256, 0).to(DEV)
    _, hp = m.encode(x)
    m.set_mode('fast')
    _, hf = m.encode(x)
    err = (hp - hf).abs().max().item()
    assert err < 2e-2, f"fast encoder differs from the fp32 encoder by {err}"
