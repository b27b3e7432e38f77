"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference leaf modules by path.

Only usable where ``/root/reference`` exists (the build container).  It is used to
  (1) pin ``oracle/reid_oracle.py`` (our CPU restatement) against the real reference code, and
  (2) generate the committed golden vectors under ``tests/golden/`` (``oracle/make_golden.py``).
Nothing in the product package, ``bench.py`` or the ``-m gpu`` tests may import this file.

The reference's ``mmdet3d/models/ReIDNet.py`` cannot be imported (needs mmcv / mmdet / pytorch3d,
SURVEY.md section 8c) but every leaf file holding the arithmetic of the path can, with three
non-invasive shims, none of which touches the reference tree:
  * ``fractions.gcd`` (removed in py3.9) is aliased to ``math.gcd`` before ``lanegcn_nets.py`` loads
    (``lanegcn_nets.py:6``);
  * the module-global ``torch`` of ``dgcnn_orig`` / ``attention`` is replaced by a proxy whose
    ``.device('cuda')`` returns the CPU device (``dgcnn_orig.py:37``, ``attention.py:115,139``);
  * a fake parent package so the relative imports of ``backbone_net.py`` resolve.
"""
import importlib.util
import math
import os
import sys
import types

REF_ROOT = os.environ.get("PCREID_REFERENCE_ROOT", "/root/reference")
_MODELS = os.path.join(REF_ROOT, "mmdet3d", "models")
_PKG = "_pcreid_ref_models"


def available():
    return os.path.isdir(_MODELS)


class _TorchProxy(types.ModuleType):
    """Forwards everything to torch but maps device('cuda') to CPU."""

    def __init__(self, real):
        super().__init__("torch")
        object.__setattr__(self, "_real", real)

    def __getattr__(self, name):
        return getattr(object.__getattribute__(self, "_real"), name)

    def device(self, *a, **k):
        real = object.__getattribute__(self, "_real")
        if a and isinstance(a[0], str) and a[0].startswith("cuda"):
            return real.device("cpu")
        return real.device(*a, **k)


def _load(name):
    full = f"{_PKG}.{name}"
    if full in sys.modules:
        return sys.modules[full]
    spec = importlib.util.spec_from_file_location(full, os.path.join(_MODELS, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Returns a namespace with the reference classes used on the hot path."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import fractions
    if not hasattr(fractions, "gcd"):
        fractions.gcd = math.gcd
    import torch
    if _PKG not in sys.modules:
        pkg = types.ModuleType(_PKG)
        pkg.__path__ = [_MODELS]
        sys.modules[_PKG] = pkg
    p2 = _load("pointnet2_utils")
    bb = _load("backbone_net")
    dg = _load("dgcnn_orig")
    pn = _load("pointnet")
    at = _load("attention")
    lg = _load("lanegcn_nets")
    proxy = _TorchProxy(torch)
    dg.torch = proxy
    at.torch = proxy
    ns = types.SimpleNamespace(
        pointnet2_utils=p2, backbone_net=bb, dgcnn_orig=dg, pointnet=pn, attention=at, lanegcn_nets=lg,
        Pointnet_Backbone=bb.Pointnet_Backbone, DGCNN=dg.DGCNN, PointNet=pn.PointNet,
        corss_attention=at.corss_attention, LinearRes=lg.LinearRes, local_self_attention=at.local_self_attention,
        cross_lin_attn=at.cross_lin_attn,
    )
    return ns
