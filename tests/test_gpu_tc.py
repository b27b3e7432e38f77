"""-m gpu: tcgen05 / TMEM primitives (csrc/tc_common.cuh) validated on hardware, one GEMM per operand configuration."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _probe_lib():
    """test-only library tests/native/libpcreid_tcprobe.so (built by __graft_entry__.build(); not part of the product)"""
    import os
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native", "libpcreid_tcprobe.so")
    assert os.path.exists(p), "run `python __graft_entry__.py` (build()) first"
    L = ctypes.CDLL(p)
    L.pcreid_tc_probe.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 4
    L.pcreid_tc_probe.restype = ctypes.c_int
    return L


def _probe(mode, N, K, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(128, K, generator=g)
    b = torch.randn(N, K, generator=g)
    if mode in (2, 4, 5):
        ad, bd = (a.to(DEV), b.to(DEV)) if mode != 4 else (a.t().contiguous().to(DEV), b.t().contiguous().to(DEV))
        # tf32 keeps 10 mantissa bits: compare against operands truncated the same way with a loose bound
        ref = a @ b.t()
        tol = 8e-3 * K ** 0.5
    else:
        a, b = (a.half(), b.half()) if mode >= 6 else (a.bfloat16(), b.bfloat16())
        ref = a.float() @ b.float().t()
        tol = 1e-4 * K ** 0.5
        if mode in (1, 7):
            ad, bd = a.t().contiguous().to(DEV), b.t().contiguous().to(DEV)
        else:
            ad, bd = a.to(DEV), b.to(DEV)
    d = torch.full((128, N), float("nan"), device=DEV)
    rc = _probe_lib().pcreid_tc_probe(mode, N, K, ctypes.c_void_p(ad.data_ptr()), ctypes.c_void_p(bd.data_ptr()),
                                    ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc == 3:
        pytest.skip("operands exceed the probe's shared-memory budget")
    assert rc == 0
    torch.cuda.synchronize()
    err = (d.cpu() - ref).abs().max().item()
    return err, tol


# mode 4 (tf32 with MN-major no-swizzle operands) is kept in the probe for reference: it does NOT produce a GEMM on sm_100a,
# which is why cn_linear_tc stages its operands K-major (mode 2)
@pytest.mark.parametrize("mode", [0, 1, 2, 3, 5, 6, 7, 8])
@pytest.mark.parametrize("N,K", [(64, 64), (128, 128), (192, 64), (64, 128), (16, 16), (256, 256), (80, 32)])
def test_tc_probe(mode, N, K):
    err, tol = _probe(mode, N, K)
    assert err < tol, f"mode {mode} N={N} K={K}: max err {err} (tol {tol})"
