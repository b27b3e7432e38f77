"""CPU checks of the torch custom-op layer (point-cloud-reid_b200/torch_ops.py): the op set is generated from include/pcreid.h,
every compute entry point of the C ABI is an op of the `pcreid` namespace with a fake kernel, the wrappers of ops/*.py,
kernels.py and models/* reach the library only through torch.ops.pcreid.*, and fake-tensor propagation works without a GPU."""
import glob
import os
import re

import pytest
import torch

from pcreid_b200 import _lib, torch_ops as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "point-cloud-reid_b200")


def test_every_compute_entry_point_is_an_op():
    structs, funcs = T.parse_header()
    assert set(funcs) == set(_lib._SIGS), set(funcs) ^ set(_lib._SIGS)          # header <-> ctypes table
    launches = {n for n, ps in funcs.items() if ps and ps[-1][1] == "stream"}
    assert {"pcreid_" + n for n in T.op_names()} == launches
    assert len(launches) >= 45
    for name in T.op_names():
        op = getattr(torch.ops.pcreid, name).default
        sp = T.SPECS[name]
        assert len(op._schema.arguments) == len(sp.flat)
        # pointer constness of the C prototype == mutation annotation of the schema
        for (kind, pname), arg in zip(sp.flat, op._schema.arguments):
            assert arg.name == pname
            assert (arg.alias_info is not None and arg.alias_info.is_write) == (kind == "M"), (name, pname)
        # ctypes signature and generated op agree on the C argument count
        n_c = sum(1 if e[0] != "struct" else 1 for e in sp.layout)
        assert n_c == len(_lib._SIGS[sp.cname]), name


def test_struct_fields_match_the_ctypes_structures():
    structs, _ = T.parse_header()
    assert [f for _, f in structs["pcreid_linear_args"]] == [f for f, _ in _lib.LinearArgs._fields_]
    assert [f for _, f in structs["pcreid_norm_args"]] == [f for f, _ in _lib.NormArgs._fields_]


def test_wrappers_reach_the_library_only_through_the_dispatcher():
    """no product module calls a compute entry point of the CDLL directly (host-side queries are allowed)"""
    host_queries = {n for n in _lib._SIGS if n[len("pcreid_"):] not in T.SPECS}
    pat = re.compile(r"\.(pcreid_\w+)\s*\(")
    for path in glob.glob(os.path.join(PKG, "**", "*.py"), recursive=True):
        if os.path.basename(path) in ("_lib.py", "torch_ops.py"):
            continue
        for m in pat.finditer(open(path).read()):
            assert m.group(1) in host_queries, (path, m.group(1))


def test_cpu_tensors_are_refused():
    x = torch.zeros(1, 8, 3)
    idx = torch.zeros(1, 4, dtype=torch.int32)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.pcreid.fps_torch(1, 8, 4, x, torch.zeros(1, dtype=torch.int32), idx)


def test_fake_tensor_propagation_without_a_gpu():
    """the public op wrappers run under FakeTensorMode on fake CUDA tensors: shapes / dtypes come out, no kernel runs"""
    from torch._subclasses.fake_tensor import FakeTensorMode
    from pcreid_b200 import kernels as K
    from pcreid_b200.ops import ball_query, furthest_point_sample, gather_points, grouping_operation, knn, three_interpolate
    calls = _lib.ABI_CALLS
    with FakeTensorMode():
        xyz = torch.empty((2, 64, 3), device="cuda")
        feats = torch.empty((2, 16, 64), device="cuda")
        fi = furthest_point_sample(xyz, 8)
        assert fi.shape == (2, 8) and fi.dtype == torch.int32 and fi.is_cuda
        centres = torch.empty((2, 8, 3), device="cuda")
        assert knn(4, xyz, centres).shape == (2, 4, 8)
        bq = ball_query(0.0, 1.0, 5, xyz, centres)
        assert bq.shape == (2, 8, 5)
        assert grouping_operation(feats, bq).shape == (2, 16, 8, 5)
        assert gather_points(feats, fi).shape == (2, 16, 8)
        w = torch.empty((2, 64, 3), device="cuda")
        assert three_interpolate(feats[:, :, :8].contiguous(), torch.empty((2, 64, 3), device="cuda", dtype=torch.int32), w).shape == (2, 16, 64)
        y = K.cn_linear(feats, torch.empty((16, 32), device="cuda"), bias=torch.empty(32, device="cuda"), act=K.ACT_RELU)
        assert y.shape == (2, 32, 64)
        assert K.knn_point(4, xyz, centres).shape == (2, 8, 4)
    assert _lib.ABI_CALLS == calls                    # nothing reached the library
