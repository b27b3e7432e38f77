"""Installs the reference's import names (`mmdet3d.ops`, `mmdet3d.models`) on top of pcreid_b200 when mmdet3d is
not importable, so that code written against the reference (`from mmdet3d.ops import knn`,
`from mmdet3d.models import build_model`) runs unchanged on the accelerated path.

If a real mmdet3d IS installed nothing of it is shadowed or replaced by default: the inference-only product class is added to
its FUSIONMODELS registry under the distinct name ``ReIDNetB200`` (configs opt in with ``type='ReIDNetB200'``).  Replacing
the registry's ``ReIDNet`` entry process-wide -- which makes every later ``build_model(dict(type='ReIDNet'))``, including a
training build, return the inference-only class -- happens only with ``install(override=True)``."""
import importlib
import logging
import sys
import types

log = logging.getLogger("pcreid_b200.compat")


def install(force=False, override=False):
    """-> True if the stand-in `mmdet3d` modules were installed, False if a real mmdet3d was found and only registered into.
    force: install the stand-in modules even if a real mmdet3d is importable.
    override: with a real mmdet3d, also replace its registry entry `ReIDNet` by the product class."""
    from . import models, ops
    real = None
    if not force:
        try:
            real = importlib.import_module("mmdet3d.models.builder")
        except ImportError:
            real = None
    if real is not None:
        real.FUSIONMODELS.register_module(name="ReIDNetB200", force=True, module=models.ReIDNet)
        log.info("real mmdet3d found: registered the accelerated inference model as FUSIONMODELS['ReIDNetB200']")
        if override:
            real.FUSIONMODELS.register_module(name="ReIDNet", force=True, module=models.ReIDNet)
            log.warning("FUSIONMODELS['ReIDNet'] now builds the inference-only pcreid_b200 model for this process "
                        "(forward_train raises); the reference class is no longer reachable through the registry")
        return False
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
