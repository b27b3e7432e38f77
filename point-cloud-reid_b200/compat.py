"""Installs the reference's import names (`mmdet3d.ops`, `mmdet3d.models`) on top of pcreid_b200 when mmdet3d is
not importable, so that code written against the reference (`from mmdet3d.ops import knn`,
`from mmdet3d.models import build_model`) runs unchanged on the accelerated path.  If a real mmdet3d is
installed, nothing is shadowed; the product classes are only registered into its FUSIONMODELS registry."""
import importlib
import sys
import types


def install(force=False):
    from . import models, ops
    try:
        if not force:
            real = importlib.import_module("mmdet3d.models.builder")
            try:
                real.FUSIONMODELS.register_module(name="ReIDNet", force=True, module=models.ReIDNet)
            except Exception:
                pass
            return False
    except Exception:
        pass
    root = types.ModuleType("mmdet3d")
    root.__path__ = []
    root.ops = ops
    root.models = models
    sys.modules["mmdet3d"] = root
    sys.modules["mmdet3d.ops"] = ops
    sys.modules["mmdet3d.models"] = models
    for sub in ("ball_query", "furthest_point_sample", "gather_points", "group_points", "knn", "interpolate", "pointnet_modules"):
        sys.modules[f"mmdet3d.ops.{sub}"] = importlib.import_module(f"{ops.__name__}.{sub}")
    return True
