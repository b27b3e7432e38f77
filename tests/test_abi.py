"""The C-ABI shared library loads and exports every symbol include/pcreid.h declares (no compute without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pcreid.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(pcreid_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from pcreid_b200 import _lib
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/pcreid.h but not exported"
    assert sorted(_lib.exported_symbols()) == names
    assert L.pcreid_abi_version() >= 1


def test_struct_layout_matches_header():
    """ctypes mirrors of pcreid_linear_args / pcreid_norm_args have the header's field order."""
    from pcreid_b200 import _lib
    src = open(os.path.join(ROOT, "include", "pcreid.h")).read()
    for struct, cls in (("pcreid_linear_args", _lib.LinearArgs), ("pcreid_norm_args", _lib.NormArgs)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), src, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.replace("*", " ").split(",")
            fields.append(names[0].split()[-1])
            fields += [n.strip() for n in names[1:]]
        assert fields == [f[0] for f in cls._fields_], struct


def test_host_helpers_without_gpu():
    from pcreid_b200 import _lib
    L = _lib.lib()
    assert [L.pcreid_fps_block_size(n) for n in (1, 2, 160, 256, 1000, 2048, 5000)] == [1, 2, 128, 256, 512, 1024, 1024]
    # argument validation happens before any CUDA call
    assert L.pcreid_knn(1, 8, 4, 101, None, None, None, None, None) == 1
    assert L.pcreid_cn_linear(None, None) == 1


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: no file of the product package (nor bench.py's measured legs) imports it, and importing
    every product module in a fresh interpreter leaves `oracle` / `tests` helpers unloaded."""
    import glob
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "point-cloud-reid_b200")
    pat = re.compile(r"^\s*(from|import)\s+(oracle|fake_kernels|helpers)\b", re.M)
    for path in glob.glob(os.path.join(pkg, "**", "*.py"), recursive=True):
        assert not pat.search(open(path).read()), path
    code = ("import sys; sys.path.insert(0, %r); import pcreid_b200, pcreid_b200.ops, pcreid_b200.models, pcreid_b200.parallel, "
            "pcreid_b200.synthetic, pcreid_b200.compat, pcreid_b200.models.tracking, pcreid_b200.models.frontend, "
            "pcreid_b200.models.image_reid; bad = [m for m in sys.modules if m == 'oracle' or m.startswith('oracle.') or "
            "m in ('fake_kernels', 'helpers')]; assert not bad, bad; print('clean')") % root
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "clean" in out.stdout, out.stderr[-1500:]
    # bench.py: oracle / tests helpers only inside the cpu_baseline / reference-arm function
    src = open(os.path.join(root, "bench.py")).read()
    body = src[src.index("def main():"):]
    assert "from oracle" not in body and "import helpers" not in body


def test_header_is_plain_c_and_cpp():
    """include/pcreid.h is the FFI contract: it must compile as C99 and as C++ with nothing but the standard headers (no torch,
    no CUDA types in the signatures)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = os.path.join(root, "include", "pcreid.h")
    for cmd in (["gcc", "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Werror", hdr], ["g++", "-fsyntax-only", "-x", "c++", "-Wall", hdr]):
        out = subprocess.run(cmd, capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
    src = open(hdr).read()
    assert "#include <torch" not in src and "#include <cuda" not in src and "at::Tensor" not in src
