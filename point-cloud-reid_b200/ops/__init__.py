"""mmdet3d.ops-compatible point ops backed by libpcreid_sm100.so (reference: mmdet3d/ops/__init__.py:11-20)."""
from . import group_points  # noqa: F401  (the reference exposes the submodule under this name too)
from .ball_query import BallQuery, ball_query
from .furthest_point_sample import (FurthestPointSampling, FurthestPointSamplingWithDist, Points_Sampler,
                                    furthest_point_sample, furthest_point_sample_with_dist)
from .gather_points import GatherPoints, gather_points
from .group_points import GroupAll, GroupingOperation, QueryAndGroup, grouping_operation
from .interpolate import ThreeInterpolate, ThreeNN, three_interpolate, three_nn
from .knn import KNN, knn
from .pointnet_modules import (SA_MODULES, ConvModule, PointFPModule, PointSAModule, PointSAModuleMSG,
                               build_sa_module)

__all__ = ["ball_query", "BallQuery", "furthest_point_sample", "furthest_point_sample_with_dist", "Points_Sampler",
           "FurthestPointSampling", "FurthestPointSamplingWithDist", "gather_points", "GatherPoints", "GroupAll",
           "QueryAndGroup", "group_points", "grouping_operation", "GroupingOperation", "knn", "KNN",
           "three_nn", "three_interpolate", "ThreeNN", "ThreeInterpolate", "PointSAModule", "PointSAModuleMSG",
           "PointFPModule", "build_sa_module", "SA_MODULES", "ConvModule"]
