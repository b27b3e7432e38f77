"""-m gpu: every fp32 row-op kernel against the torch emulation of its specification (tests/fake_kernels.py)."""
import pytest
import torch

import fake_kernels as F
import pcreid_b200.kernels as K

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def close(a, b, tol=2e-5):
    a, b = a.cpu(), b.cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item()
    scale = max(1.0, b.abs().max().item())
    assert err <= tol * scale, f"max err {err} (scale {scale})"


@pytest.mark.parametrize("B,K1,CO,N", [(3, 64, 64, 256), (2, 3, 32, 160), (2, 128, 192, 130), (1, 1024, 512, 77), (5, 67, 9, 33)])
@pytest.mark.parametrize("act", [0, 1, 2])
def test_cn_linear_basic(B, K1, CO, N, act):
    x, w, b = rnd(B, K1, N, seed=1), rnd(K1, CO, seed=2) / K1 ** 0.5, rnd(CO, seed=3)
    close(K.cn_linear(x.to(DEV), w.to(DEV), bias=b.to(DEV), act=act), F.cn_linear(x, w, bias=b, act=act))


def test_cn_linear_two_inputs_residual_maps_strides():
    B, N = 6, 200
    x1, x2 = rnd(4, 64, N, seed=1), rnd(3, N, 3, seed=2)
    w1, w2 = rnd(64, 128, seed=3) / 8, rnd(3, 128, seed=4)
    res = rnd(5, 128, N, seed=5)
    m1 = torch.tensor([0, 3, 1, 1, 2, 0], dtype=torch.int32)
    m2 = torch.tensor([2, 2, 0, 1, 1, 0], dtype=torch.int32)
    mr = torch.tensor([4, 0, 1, 2, 3, 3], dtype=torch.int32)
    for after in (False, True):
        got = K.cn_linear(x1.to(DEV), w1.to(DEV), x2=x2.to(DEV), w2=w2.to(DEV), x2_pm=True, act=1, res=res.to(DEV),
                          res_after_act=after, x1_map=m1.to(DEV), x2_map=m2.to(DEV), r_map=mr.to(DEV), B=B)
        close(got, F.cn_linear(x1, w1, x2=x2, w2=w2, x2_pm=True, act=1, res=res, res_after_act=after, x1_map=m1, x2_map=m2,
                               r_map=mr, B=B))
    # strided views (qkv slices), rows < N, per-object weights with a map, point-major output
    qkv = rnd(3, 192, N, seed=6)
    wk = rnd(4, 64, 64, seed=7) / 8
    wm = torch.tensor([3, 0, 2], dtype=torch.int32)
    qd = qkv.to(DEV)
    close(K.cn_linear(qd[:, 64:128], wk.to(DEV), w1_map=wm.to(DEV), rows=96), F.cn_linear(qkv[:, 64:128], wk, w1_map=wm, rows=96))
    close(K.cn_linear(qd[:, 128:], wk[0].contiguous().to(DEV), y_pm=True), F.cn_linear(qkv[:, 128:], wk[0], y_pm=True))
    out = torch.zeros(3, 256, N, device=DEV)
    K.cn_linear(qd[:, :64], wk[1].contiguous().to(DEV), out=out[:, 64:128])
    close(out[:, 64:128], F.cn_linear(qkv[:, :64], wk[1]))
    assert float(out[:, :64].abs().max()) == 0 and float(out[:, 128:].abs().max()) == 0


@pytest.mark.parametrize("B,K1,CO,N", [(3, 64, 64, 256), (2, 128, 192, 132), (2, 1024, 512, 256), (5, 72, 12, 36), (1, 512, 1024, 100)])
@pytest.mark.parametrize("act", [0, 2])
@pytest.mark.parametrize("gen", [1, 2])
def test_cn_linear_tensor_core(B, K1, CO, N, act, gen):
    """tcgen05 kind::tf32 GEMMs (gen 1: cn_linear_tc.cu, gen 2: warp-specialised cn_linear_tc2.cu) against the fp32
    specification (tf32 operands: 10-bit mantissa)."""
    x, w, b = rnd(B, K1, N, seed=1), rnd(K1, CO, seed=2) / K1 ** 0.5, rnd(CO, seed=3)
    K._TC_LINEAR["min_k"] = 8            # force the tensor-core kernel for every shape under test
    K._TC_LINEAR["gen"] = gen
    K._TC_LINEAR["tma"] = False
    try:
        with K.tensor_core_linear(True):
            got = K.cn_linear(x.to(DEV), w.to(DEV), bias=b.to(DEV), act=act)
    finally:
        K._TC_LINEAR.pop("min_k")
        K._TC_LINEAR.pop("gen")
        K._TC_LINEAR.pop("tma")
    close(got, F.cn_linear(x, w, bias=b, act=act), 3e-3)


@pytest.fixture
def tma_only():
    """cn_linear must be served by pcreid_cn_linear_tma: the older tensor-core generations and the FFMA kernel are made to fail"""
    K._TC_LINEAR["tma"] = True
    real_tc, real_tc2, real_ffma = K._OPS.cn_linear_tc, K._OPS.cn_linear_tc2, K._OPS.cn_linear
    calls = {"n": 0}
    real_tma = K._OPS.cn_linear_tma

    class _Ops:
        def __getattr__(self, name):
            if name in ("cn_linear_tc", "cn_linear_tc2", "cn_linear"):
                raise AssertionError(f"{name} used although the TMA kernel supports the shape")
            if name == "cn_linear_tma":
                def counted(*a):
                    calls["n"] += 1
                    return real_tma(*a)
                return counted
            return getattr(real_ops, name)
    real_ops = K._OPS
    K._OPS = _Ops()
    try:
        with K.tensor_core_linear(True):
            yield calls
    finally:
        K._OPS = real_ops
        K._TC_LINEAR.pop("tma")


@pytest.mark.parametrize("B,K1,CO,N", [(3, 64, 64, 256), (2, 128, 192, 132), (2, 1024, 512, 256), (5, 72, 36, 36), (1, 512, 1024, 100),
                                       (4, 3, 64, 128), (2, 6, 32, 20), (3, 40, 100, 260), (2, 96, 160, 1024)])
@pytest.mark.parametrize("act", [0, 2])
@pytest.mark.parametrize("tf32_maps", [True, False])
def test_cn_linear_tma(B, K1, CO, N, act, tf32_maps, tma_only):
    """TMA tensor-map staged tcgen05 kind::tf32 GEMM (cn_linear_tma.cu): K tails / K = 3 zero-filled by the TMA unit, ragged point
    and channel tiles, against the fp32 specification."""
    x, w, b = rnd(B, K1, N, seed=1), rnd(K1, CO, seed=2) / K1 ** 0.5, rnd(CO, seed=3)
    K._TC_LINEAR["tma_tf32_maps"] = tf32_maps
    try:
        got = K.cn_linear(x.to(DEV), w.to(DEV), bias=b.to(DEV), act=act)
    finally:
        K._TC_LINEAR.pop("tma_tf32_maps")
    assert tma_only["n"] == 1
    close(got, F.cn_linear(x, w, bias=b, act=act), 3e-3)


def test_cn_linear_tma_options(tma_only):
    """second input pair, residual before / after the activation, object maps on every operand, per-object weights, strided
    channel views, rows < N, point-major output, output into a channel slice"""
    N, B = 200, 6
    x1, x2, res = rnd(4, 64, N, seed=1), rnd(5, 32, N, seed=2), rnd(3, 128, N, seed=5)
    w1, w2 = rnd(64, 128, seed=3) / 8, rnd(32, 128, seed=4) / 6
    m1 = torch.tensor([3, 0, 2, 2, 1, 0], dtype=torch.int32)
    m2 = torch.tensor([4, 4, 0, 1, 3, 2], dtype=torch.int32)
    mr = torch.tensor([0, 2, 1, 1, 2, 0], dtype=torch.int32)
    for after in (False, True):
        got = K.cn_linear(x1.to(DEV), w1.to(DEV), x2=x2.to(DEV), w2=w2.to(DEV), act=1, res=res.to(DEV), res_after_act=after,
                          x1_map=m1.to(DEV), x2_map=m2.to(DEV), r_map=mr.to(DEV), B=B)
        close(got, F.cn_linear(x1, w1, x2=x2, w2=w2, act=1, res=res, res_after_act=after, x1_map=m1, x2_map=m2, r_map=mr, B=B), 3e-3)
    wk = rnd(4, 64, 64, seed=7) / 8                          # per-object weights
    close(K.cn_linear(x1.to(DEV), wk.to(DEV)), F.cn_linear(x1, wk), 3e-3)
    wm = torch.tensor([3, 0, 2], dtype=torch.int32)
    qkv = rnd(3, 192, N, seed=6)
    qd = qkv.to(DEV)
    close(K.cn_linear(qd[:, 64:128], wk.to(DEV), w1_map=wm.to(DEV), rows=96), F.cn_linear(qkv[:, 64:128], wk, w1_map=wm, rows=96), 3e-3)
    out = torch.zeros(3, 256, N, device=DEV)
    K.cn_linear(qd[:, :64], wk[1].contiguous().to(DEV), out=out[:, 64:128])
    close(out[:, 64:128], F.cn_linear(qkv[:, :64], wk[1]), 3e-3)
    assert float(out[:, :64].abs().max()) == 0 and float(out[:, 128:].abs().max()) == 0
    # ragged rows next to foreign data: the tensor store clips at `rows`, columns [rows, N) of the output keep their content
    out = torch.full((3, 64, N), 7.0, device=DEV)
    K.cn_linear(qd[:, :64], wk[1].contiguous().to(DEV), out=out, rows=92)
    close(out[:, :, :92], F.cn_linear(qkv[:, :64], wk[1], rows=92), 3e-3)
    assert bool((out[:, :, 92:] == 7.0).all())


def test_cn_linear_tma_unsupported_shapes_fall_back():
    """point-major outputs / inputs and CO < 32 are answered PCREID_ERR_UNSUPPORTED and served by the older kernels"""
    N = 200
    qkv, wk = rnd(3, 192, N, seed=6), rnd(64, 64, seed=7) / 8
    K._TC_LINEAR["tma"] = True
    try:
        with K.tensor_core_linear(True):
            close(K.cn_linear(qkv.to(DEV)[:, 128:], wk.to(DEV), y_pm=True), F.cn_linear(qkv[:, 128:], wk, y_pm=True), 3e-3)
            x, w = rnd(2, 64, N, seed=1), rnd(64, 12, seed=2) / 8
            close(K.cn_linear(x.to(DEV), w.to(DEV)), F.cn_linear(x, w), 3e-3)
            # the TMA unit clips the innermost extent in 16-byte units: row counts that are not multiples of 4 stay on the older kernels
            out = torch.full((3, 64, N), 7.0, device=DEV)
            K.cn_linear(qkv.to(DEV)[:, :64], wk.to(DEV), out=out, rows=90)
            close(out[:, :, :90], F.cn_linear(qkv[:, :64], wk, rows=90), 3e-3)
            assert bool((out[:, :, 90:] == 7.0).all())
    finally:
        K._TC_LINEAR.pop("tma")


def test_cn_linear_tma_many_tiles_per_cta(tma_only):
    """persistent loop: several tiles per CTA, both accumulator buffers and every ring slot reused many times"""
    B, K1, K2, CO, N = 700, 256, 64, 320, 200
    x1, x2 = rnd(B, K1, N, seed=1), rnd(B, K2, N, seed=2)
    w1, w2, b = rnd(K1, CO, seed=3) / 16, rnd(K2, CO, seed=4) / 8, rnd(CO, seed=5)
    res = rnd(B, CO, N, seed=6)
    got = K.cn_linear(x1.to(DEV), w1.to(DEV), x2=x2.to(DEV), w2=w2.to(DEV), bias=b.to(DEV), act=1, res=res.to(DEV))
    close(got, F.cn_linear(x1, w1, x2=x2, w2=w2, bias=b, act=1, res=res), 3e-3)


@pytest.mark.parametrize("B,K1,CO,N", [(3, 64, 64, 256), (2, 128, 192, 132), (2, 1024, 512, 256), (4, 3, 64, 128), (3, 40, 100, 260)])
@pytest.mark.parametrize("act", [0, 1])
def test_cn_linear_tma_x3_is_fp32_grade(B, K1, CO, N, act, monkeypatch):
    monkeypatch.setattr(K, "X3_MAX_K", 1 << 20)          # the dispatch keeps K > 256 on the FFMA kernel; the kernel itself is tested beyond
    """3 x tf32 (hi + lo operand split, pcreid_cn_linear_tma_x3) against a float64 reference: within 1e-5 of the output scale up to
    K = 1024 (measured 2e-6 at K <= 128, 8e-6 at K = 1024: the tensor core's fp32 accumulation truncates; the FFMA kernel sits at
    1.5e-6, single-pass tf32 at 1.5e-4), operands with full 24-bit significands"""
    x, w, b = rnd(B, K1, N, seed=1) * 3.7, rnd(K1, CO, seed=2) / K1 ** 0.5, rnd(CO, seed=3)
    ref = torch.einsum("bkn,kc->bcn", x.double(), w.double()) + b.double()[None, :, None]
    if act:
        ref = ref.clamp_min(0)
    calls = {"n": 0}
    real = K._OPS.cn_linear_tma_x3

    class _Ops:
        def __getattr__(self, name):
            if name == "cn_linear_tma_x3":
                def f(*a):
                    calls["n"] += 1
                    return real(*a)
                return f
            if name in ("cn_linear", "cn_linear_tc", "cn_linear_tc2", "cn_linear_tma"):
                raise AssertionError(f"{name} used instead of the 3 x tf32 kernel")
            return getattr(ops, name)
    ops = K._OPS
    K._OPS = _Ops()
    try:
        with K.tensor_core_linear(True, min_k=1 << 30, x3=True):
            got = K.cn_linear(x.to(DEV), w.to(DEV), bias=b.to(DEV), act=act)
    finally:
        K._OPS = ops
    assert calls["n"] == 1
    err = float((got.cpu().double() - ref).abs().max())
    ffma = float((K.cn_linear(x.to(DEV), w.to(DEV), bias=b.to(DEV), act=act).cpu().double() - ref).abs().max())
    scale = max(1.0, float(ref.abs().max()))
    assert err <= 1e-5 * scale, (err, ffma, scale)


def test_cn_linear_tma_x3_options():
    """second input pair, residual, object maps on every operand, per-object weights with a map (their lo part is split per call)"""
    N, B = 200, 6
    x1, x2, res = rnd(4, 64, N, seed=1), rnd(5, 32, N, seed=2), rnd(3, 128, N, seed=5)
    w1, w2 = rnd(64, 128, seed=3) / 8, rnd(32, 128, seed=4) / 6
    m1 = torch.tensor([3, 0, 2, 2, 1, 0], dtype=torch.int32)
    m2 = torch.tensor([4, 4, 0, 1, 3, 2], dtype=torch.int32)
    mr = torch.tensor([0, 2, 1, 1, 2, 0], dtype=torch.int32)
    wk = rnd(4, 64, 64, seed=7) / 8
    wm = torch.tensor([3, 0, 2], dtype=torch.int32)
    qkv = rnd(3, 192, N, seed=6)
    with K.tensor_core_linear(True, min_k=1 << 30, x3=True):
        got = K.cn_linear(x1.to(DEV), w1.to(DEV), x2=x2.to(DEV), w2=w2.to(DEV), act=1, res=res.to(DEV), res_after_act=True,
                          x1_map=m1.to(DEV), x2_map=m2.to(DEV), r_map=mr.to(DEV), B=B)
        close(got, F.cn_linear(x1, w1, x2=x2, w2=w2, act=1, res=res, res_after_act=True, x1_map=m1, x2_map=m2, r_map=mr, B=B))
        close(K.cn_linear(qkv.to(DEV)[:, 64:128], wk.to(DEV), w1_map=wm.to(DEV), rows=96), F.cn_linear(qkv[:, 64:128], wk, w1_map=wm, rows=96))


def test_cn_linear_tma_rounding():
    """what the two tensor-map element types do with the 13 low mantissa bits of an fp32 activation: FLOAT32 maps leave them to the
    tensor core (which ignores them: truncation), TFLOAT32 maps round to nearest in the TMA unit; ROUND_OUT rounds the result."""
    N = 128
    x = torch.full((1, 32, N), 1.0 + 2.0 ** -11 + 2.0 ** -13)            # tf32: truncates to 1, rounds to 1 + 2^-10
    w = torch.zeros(32, 32)
    w[0, 0] = 1.0
    K._TC_LINEAR["tma"] = True
    try:
        with K.tensor_core_linear(True):
            res = {}
            for tf32_maps in (False, True):
                K._TC_LINEAR["tma_tf32_maps"] = tf32_maps
                res[tf32_maps] = float(K.cn_linear(x.to(DEV), w.to(DEV))[0, 0, 5])
            K._TC_LINEAR["round_out"] = True
            y = K.cn_linear((x * 3).to(DEV), w.to(DEV))
    finally:
        for k in ("tma", "tma_tf32_maps", "round_out"):
            K._TC_LINEAR.pop(k, None)
    assert res[False] == 1.0, res
    assert res[True] == 1.0 + 2.0 ** -10, res
    assert torch.equal(y.view(torch.int32) & 0x1fff, torch.zeros_like(y, dtype=torch.int32))


def test_cn_linear_tc2_many_tiles_per_cta():
    """persistent loop: several tiles per CTA, both accumulator buffers and every ring slot reused many times"""
    B, K1, K2, CO, N = 700, 256, 64, 320, 200
    x1, x2 = rnd(B, K1, N, seed=1), rnd(B, K2, N, seed=2)
    w1, w2, b = rnd(K1, CO, seed=3) / 16, rnd(K2, CO, seed=4) / 8, rnd(CO, seed=5)
    res = rnd(B, CO, N, seed=6)
    K._TC_LINEAR["tma"] = False
    try:
        with K.tensor_core_linear(True):
            got = K.cn_linear(x1.to(DEV), w1.to(DEV), x2=x2.to(DEV), w2=w2.to(DEV), bias=b.to(DEV), act=1, res=res.to(DEV))
    finally:
        K._TC_LINEAR.pop("tma")
    close(got, F.cn_linear(x1, w1, x2=x2, w2=w2, bias=b, act=1, res=res), 3e-3)


def test_cn_linear_tensor_core_options():
    N = 200
    x1, x2, res = rnd(4, 64, N, seed=1), rnd(4, 32, N, seed=2), rnd(4, 128, N, seed=5)
    w1, w2 = rnd(64, 128, seed=3) / 8, rnd(32, 128, seed=4) / 6
    wk = rnd(4, 64, 64, seed=7) / 8                          # per-object weights
    qkv = rnd(3, 192, N, seed=6)
    K._TC_LINEAR["min_k"] = 8
    K._TC_LINEAR["tma"] = False
    with K.tensor_core_linear(True):
        for after in (False, True):
            got = K.cn_linear(x1.to(DEV), w1.to(DEV), x2=x2.to(DEV), w2=w2.to(DEV), act=1, res=res.to(DEV), res_after_act=after)
            close(got, F.cn_linear(x1, w1, x2=x2, w2=w2, act=1, res=res, res_after_act=after), 3e-3)
        close(K.cn_linear(x1.to(DEV), wk.to(DEV)), F.cn_linear(x1, wk), 3e-3)
        qd = qkv.to(DEV)
        close(K.cn_linear(qd[:, 64:128], wk[0].contiguous().to(DEV), rows=96), F.cn_linear(qkv[:, 64:128], wk[0], rows=96), 3e-3)
        close(K.cn_linear(qd[:, 128:], wk[1].contiguous().to(DEV), y_pm=True), F.cn_linear(qkv[:, 128:], wk[1], y_pm=True), 3e-3)
        # shapes outside the tensor-core kernel silently use the FFMA kernel (exact to 2e-5)
        xyz = rnd(2, N, 3, seed=9)
        w3 = rnd(3, 32, seed=10)
        close(K.cn_linear(xyz.to(DEV), w3.to(DEV), x1_pm=True), F.cn_linear(xyz, w3, x1_pm=True))
    K._TC_LINEAR.pop("min_k")
    K._TC_LINEAR.pop("tma")


@pytest.mark.parametrize("C,G,N", [(64, 1, 256), (128, 8, 100), (512, 64, 33), (32, 1, 7)])
def test_cn_groupnorm(C, G, N):
    x, g, b, r = rnd(3, C, N, seed=1) * 3 + 1, rnd(C, seed=2), rnd(C, seed=3), rnd(2, C, N, seed=4)
    rm = torch.tensor([1, 0, 1], dtype=torch.int32)
    close(K.cn_groupnorm(x.to(DEV), g.to(DEV), b.to(DEV), G), F.cn_groupnorm(x, g, b, G), 1e-4)
    close(K.cn_groupnorm(x.to(DEV), g.to(DEV), b.to(DEV), G, res=r.to(DEV), r_map=rm.to(DEV), act=1),
          F.cn_groupnorm(x, g, b, G, res=r, r_map=rm, act=1), 1e-4)


@pytest.mark.parametrize("d,H,S", [(64, 2, 256), (32, 2, 100), (128, 2, 300), (64, 2, 37)])
def test_linear_attention_pieces(d, H, S):
    k, v, q = rnd(3, d, S, seed=1), rnd(3, d, S, seed=2), rnd(3, d, 150, seed=3)
    wkv, ks = K.linattn_kv(k.to(DEV), v.to(DEV), H)
    wkv_r, ks_r = F.linattn_kv(k, v, H)
    close(wkv, wkv_r, 1e-5)
    close(ks, ks_r, 1e-5)
    qm = torch.tensor([2, 0, 0, 1], dtype=torch.int32)
    km = torch.tensor([1, 1, 2, 0], dtype=torch.int32)
    close(K.linattn_scale(q.to(DEV), ks, H, S, q_map=qm.to(DEV), ksum_map=km.to(DEV)),
          F.linattn_scale(q, ks_r, H, S, q_map=qm, ksum_map=km), 1e-4)


def test_pooling():
    a, b = rnd(5, 64, 256, seed=1), rnd(5, 64, 200, seed=2)
    close(K.cn_pool(a.to(DEV), b.to(DEV), mode=0), F.cn_pool(a, b, mode=0))
    close(K.cn_pool(a.to(DEV), mode=1, transposed=True), F.cn_pool(a, mode=1, transposed=True))
    close(K.cn_chanmax(a.to(DEV)), F.cn_chanmax(a))
    close(K.cn_chanmax(a.to(DEV), transposed=True), F.cn_chanmax(a, transposed=True))


@pytest.mark.parametrize("C,N,S,k", [(32, 256, 256, 32), (64, 256, 128, 48), (128, 128, 64, 48), (64, 160, 80, 48), (32, 40, 37, 20)])
def test_sa_edge_mlp(C, N, S, k):
    g = torch.Generator().manual_seed(C + k)
    p1, cc = rnd(3, C, N, seed=1), rnd(3, C, S, seed=2)
    idx = torch.randint(0, N, (3, S, k), generator=g, dtype=torch.int32)
    w2, w3 = rnd(C, C, seed=3) / C ** 0.5, rnd(C, C, seed=4) / C ** 0.5
    b2, b3 = rnd(C, seed=5) * 0.1, rnd(C, seed=6) * 0.1
    got = K.sa_edge_mlp(p1.to(DEV), cc.to(DEV), idx.to(DEV), w2.to(DEV), b2.to(DEV), w3.to(DEV), b3.to(DEV))
    close(got, F.sa_edge_mlp(p1, cc, idx, w2, b2, w3, b3))


@pytest.mark.parametrize("gen", [1, 2])
@pytest.mark.parametrize("C,N,S,k", [(32, 256, 256, 32), (64, 256, 128, 48), (128, 128, 64, 48), (64, 160, 80, 48), (32, 40, 37, 20),
                                     (128, 1024, 512, 48), (64, 64, 33, 16), (32, 128, 50, 63), (128, 256, 7, 40)])
def test_sa_edge_mlp_tensor_core(C, N, S, k, gen):
    """tcgen05 kind::tf32 version against the fp32 specification: tf32 operands (10-bit mantissa) -> 5e-3 of the scale."""
    g = torch.Generator().manual_seed(C + k)
    B = 3 if N < 1024 else 160        # the large case spans several persistent waves
    p1, cc = rnd(B, C, N, seed=1), rnd(B, C, S, seed=2)
    idx = torch.randint(0, N, (B, S, k), generator=g, dtype=torch.int32)
    w2, w3 = rnd(C, C, seed=3) / C ** 0.5, rnd(C, C, seed=4) / C ** 0.5      # (C_out, C_in)
    b2, b3 = rnd(C, seed=5) * 0.1, rnd(C, seed=6) * 0.1
    got = K.sa_edge_mlp_tc(p1.transpose(1, 2).contiguous().to(DEV), cc.transpose(1, 2).contiguous().to(DEV), idx.to(DEV),
                           K.tf32_image(w2).to(DEV), b2.to(DEV), K.tf32_image(w3).to(DEV), b3.to(DEV), gen=gen)
    close(got, F.sa_edge_mlp(p1, cc, idx, w2.t().contiguous(), b2, w3.t().contiguous(), b3), 5e-3)


def test_edge_gather_max_into_strided_output():
    g = torch.Generator().manual_seed(3)
    p, q = rnd(2, 64, 300, seed=1), rnd(2, 64, 300, seed=2)
    idx = torch.randint(0, 300, (2, 300, 20), generator=g, dtype=torch.int32)
    cat = torch.zeros(2, 512, 300, device=DEV)
    K.edge_gather_max(p.to(DEV), q.to(DEV), idx.to(DEV), 2, out=cat[:, 64:128])
    close(cat[:, 64:128], F.edge_gather_max(p, q, idx, 2))


@pytest.mark.parametrize("T,D", [(5, 70), (33, 32)])
def test_pair_concat_head(T, D):
    E, G = 128, 32
    a, bv, et, ed = rnd(T, 2 * E, seed=1), rnd(D, 2 * E, seed=2), rnd(T, E, seed=3), rnd(D, E, seed=4)
    w2 = rnd(2 * E, 2 * E, seed=5) / 16
    g1, b1, g2, b2, w = rnd(2 * E, seed=6), rnd(2 * E, seed=7), rnd(2 * E, seed=8), rnd(2 * E, seed=9), rnd(2 * E, seed=10) / 16
    mask = (torch.rand(T, D, generator=torch.Generator().manual_seed(1)) > 0.3).to(torch.uint8)
    dv = lambda *ts: [t.to(DEV) for t in ts]
    got = K.pair_concat_head(*dv(a, bv, et, ed, w2, g1, b1, g2, b2, w), 0.25, G, mask.to(DEV))
    close(got, F.pair_concat_head(a, bv, et, ed, w2, g1, b1, g2, b2, w, 0.25, G, mask), 1e-4)


@pytest.mark.parametrize("C,G,P", [(128, 8, 1000), (128, 16, 65), (256, 32, 300), (96, 8, 17)])
def test_gn_res_relu_dot_matches_the_unfused_head_tail(C, G, P):
    """pcreid_gn_res_relu_dot = cn_groupnorm(.., res, ReLU) followed by the Linear(C, 1): same GroupNorm arithmetic, the dot as one
    fma chain over the channels (the unfused GEMM sums in the same order: 1e-6 of the scale)."""
    x, r = rnd(1, C, P, seed=1), rnd(1, C, P, seed=2)
    g, b, w = rnd(C, seed=3), rnd(C, seed=4), rnd(C, seed=5) / C ** 0.5
    got = K.gn_res_relu_dot(x.to(DEV), g.to(DEV), b.to(DEV), G, r.to(DEV), w.to(DEV), 0.25)
    y = K.cn_groupnorm(x.to(DEV), g.to(DEV), b.to(DEV), G, res=r.to(DEV), act=K.ACT_RELU)
    want = (y[0] * w.to(DEV)[:, None]).sum(0) + 0.25
    assert got.shape == (P,)
    close(got, want, 1e-5)
    xg = x.view(1, G, C // G, P)
    ref = ((xg - xg.mean(2, keepdim=True)) / torch.sqrt(xg.var(2, unbiased=False, keepdim=True) + 1e-5)).view(1, C, P)
    ref = torch.relu(ref * g[None, :, None] + b[None, :, None] + r)[0]
    close(got, (ref * w[:, None]).sum(0) + 0.25, 1e-4)
