"""`torch.ops.pcreid.*` -- the torch custom-op layer over the C ABI of libpcreid_sm100.so.

The reference reaches its CUDA launchers through pybind11 modules, one `*_wrapper(sizes..., at::Tensor...)` per launcher
(mmdet3d/ops/knn/src/knn.cpp:15-40, ops/ball_query/src/ball_query.cpp:20-42, ops/furthest_point_sample/src/
furthest_point_sample.cpp:24-60, ops/group_points/src/group_points.cpp:28-58, ops/gather_points/src/gather_points.cpp:23-50):
sizes as ints, tensors in, preallocated result tensors filled in place.  This module is that layer for the drop-in, as
dispatcher ops instead of pybind functions: ONE op per compute entry point of include/pcreid.h, generated from the header
itself, so the op set cannot drift from the ABI:

    int pcreid_knn_t(int b, int n, int m, int nsample, const float* xyz, const float* new_xyz, int* idx, float* dist2, void* stream)
 -> pcreid::knn_t(int b, int n, int m, int nsample, Tensor? xyz, Tensor? new_xyz, Tensor(a!)? idx, Tensor(b!)? dist2) -> ()

  * `const T*`  -> `Tensor?`  (input; its data pointer is passed; None -> NULL)
  * `T*`        -> `Tensor(x!)?` (result / scratch buffer written by the kernel: declared as mutated; None -> NULL)
  * `int` / `long long` / `float` -> `int` / `int` / `float`
  * `void* stream` is not an op argument: the CUDA implementation passes the current stream of the tensors' device
  * `const pcreid_linear_args*` / `const pcreid_norm_args*` are flattened into their fields (same rules)
  * host-side queries (no stream argument: pcreid_abi_version, *_blob_bytes, ...) are not kernels and stay plain ctypes calls

Registered per op: the CUDA implementation (ctypes call into the library; any other backend fails in the dispatcher --
there is no CPU kernel) and a fake / meta implementation (the ops return nothing: outputs are preallocated by the callers
in ops/*.py, kernels.py and models/fused_pairs.py, so shape inference is theirs), which is what FakeTensorMode, opcheck and
torch.compile need.  A non-zero return code raises; the three entry points that answer PCREID_ERR_UNSUPPORTED for shapes
outside their tiles (`cn_linear_tc`, `cn_linear_tc2`, `cn_linear_tma`, `cn_linear_tma_x3`, `sa_edge_mlp_tc2`) return the code instead (`-> int`).
"""
import ctypes
import os
import re

import torch

from . import _lib

NAMESPACE = "pcreid"
_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "pcreid.h")
RC_OPS = ("cn_linear_tc", "cn_linear_tc2", "cn_linear_tma", "cn_linear_tma_x3", "sa_edge_mlp_tc2")      # return the code (UNSUPPORTED is an answer, not an error)

_STRUCTS = {"pcreid_linear_args": _lib.LinearArgs, "pcreid_norm_args": _lib.NormArgs}


def _strip_comments(src):
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", " ", src)


def _kind(ctype, name):
    """C parameter -> (kind, schema type): kind in 'i' (int), 'f' (float), 'T' (input tensor), 'M' (mutated tensor), 's' (stream)."""
    if "*" in ctype:
        if name == "stream":
            return "s"
        return "T" if "const" in ctype else "M"
    if "float" in ctype or "double" in ctype:
        return "f"
    return "i"


def _params(arglist):
    out = []
    for a in arglist.split(","):
        a = a.strip()
        if not a or a == "void":
            continue
        m = re.match(r"^(.*?)(\w+)$", a)
        out.append((m.group(1).strip(), m.group(2)))
    return out


def parse_header(path=HEADER):
    """-> ({struct name: [(ctype, field)]}, {function name: [(ctype, param)]}) of include/pcreid.h."""
    src = _strip_comments(open(path).read())
    structs = {}
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*\w+\s*;", src, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            mm = re.match(r"^(.*?)(\w+(?:\s*,\s*\w+)*)$", decl, flags=re.S)
            ctype = mm.group(1).strip()
            for nm in mm.group(2).split(","):
                fields.append((ctype, nm.strip()))
        structs[m.group(1)] = fields
    funcs = {}
    for m in re.finditer(r"\bint\s+(pcreid_\w+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        funcs[m.group(1)] = _params(m.group(2))
    return structs, funcs


class OpSpec:
    """one generated op: flat argument list [(kind, name)], how they map back onto the C call."""
    __slots__ = ("name", "cname", "flat", "layout", "returns_rc", "schema")


def _build_specs():
    structs, funcs = parse_header()
    specs = {}
    for cname, params in funcs.items():
        if not params or params[-1][1] != "stream":
            continue                                   # host-side query, not a kernel launch
        sp = OpSpec()
        sp.cname, sp.name = cname, cname[len("pcreid_"):]
        sp.flat, sp.layout = [], []                    # layout: ('arg', flat index) | ('struct', struct name, [flat indices]) | ('stream',)
        for ctype, pname in params:
            sname = next((s for s in structs if s in ctype), None)
            if sname is not None:
                idxs = []
                for ft, fn in structs[sname]:
                    idxs.append(len(sp.flat))
                    sp.flat.append((_kind(ft, fn), fn))
                sp.layout.append(("struct", sname, idxs))
            elif _kind(ctype, pname) == "s":
                sp.layout.append(("stream",))
            else:
                sp.layout.append(("arg", len(sp.flat)))
                sp.flat.append((_kind(ctype, pname), pname))
        sp.returns_rc = sp.name in RC_OPS
        parts, alias = [], iter("abcdefghijklmnopqrstuvwxyz")
        for kind, pname in sp.flat:
            ty = {"i": "int", "f": "float", "T": "Tensor?"}.get(kind) or f"Tensor({next(alias)}!)?"
            parts.append(f"{ty} {pname}")
        sp.schema = f"{sp.name}({', '.join(parts)}) -> {'int' if sp.returns_rc else '()'}"
        specs[sp.name] = sp
    return specs, structs


SPECS, _STRUCT_FIELDS = _build_specs()
_LIBRARY = torch.library.Library(NAMESPACE, "DEF")


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _make_cuda_impl(sp):
    kinds = [k for k, _ in sp.flat]
    tensor_pos = [i for i, k in enumerate(kinds) if k in "TM"]
    name, cname, layout, returns_rc = sp.name, sp.cname, sp.layout, sp.returns_rc

    def impl(*args):
        dev = None
        for i in tensor_pos:
            t = args[i]
            if t is not None:
                if not t.is_cuda:
                    raise RuntimeError(f"pcreid::{name}: every tensor must be a CUDA tensor (there is no CPU kernel)")
                dev = t.device if dev is None else dev
        cargs, keep = [], []
        for ent in layout:
            if ent[0] == "arg":
                a = args[ent[1]]
                cargs.append(_ptr(a) if kinds[ent[1]] in "TM" else a)
            elif ent[0] == "struct":
                st = _STRUCTS[ent[1]]()
                for (ft, fn), i in zip(_STRUCT_FIELDS[ent[1]], ent[2]):
                    v = args[i]
                    setattr(st, fn, (_ptr(v) if kinds[i] in "TM" else v))
                keep.append(st)
                cargs.append(ctypes.byref(st))
            else:
                cargs.append(ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        fn = getattr(_lib.lib(), cname)
        if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device():
            with torch.cuda.device(dev):
                rc = fn(*cargs)
        else:
            rc = fn(*cargs)
        if returns_rc and rc == 3:
            _lib.ABI_CALLS += 1
            return rc
        _lib.check(rc, cname)
        return rc if returns_rc else None

    return impl


def _make_fake_impl(sp):
    if sp.returns_rc:
        return lambda *args: 0
    return lambda *args: None


for _sp in SPECS.values():
    _LIBRARY.define(_sp.schema)
    _LIBRARY.impl(_sp.name, _make_cuda_impl(_sp), "CUDA")
    torch.library.register_fake(f"{NAMESPACE}::{_sp.name}", _make_fake_impl(_sp), lib=_LIBRARY)

ops = getattr(torch.ops, NAMESPACE)


def op_names():
    return sorted(SPECS)


class StructArgs:
    """keyword view of a flattened struct argument block: fields default to 0 / None, tensors are stored as tensors."""

    def __init__(self, sname):
        object.__setattr__(self, "_fields", [fn for _, fn in _STRUCT_FIELDS[sname]])
        object.__setattr__(self, "_kinds", {fn: _kind(ft, fn) for ft, fn in _STRUCT_FIELDS[sname]})
        for fn in self._fields:
            object.__setattr__(self, fn, None if self._kinds[fn] in "TM" else 0)

    def __setattr__(self, k, v):
        if k not in self._kinds:
            raise AttributeError(k)
        object.__setattr__(self, k, v)

    def astuple(self):
        return tuple(getattr(self, fn) for fn in self._fields)


def linear_args():
    return StructArgs("pcreid_linear_args")


def norm_args():
    return StructArgs("pcreid_norm_args")
