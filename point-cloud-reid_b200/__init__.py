"""pcreid-b200: B200-native (sm_100a) implementation of the point-set encoder + siamese match-head hot
path of bentherien/point-cloud-reid, behind the reference's model / op API.

Sub-packages
  ops     -- mmdet3d.ops-compatible point ops (furthest_point_sample, knn, ball_query, group_points, ...)
  models  -- ReIDNet and its backbones / attention blocks / heads with the reference's names,
             constructor kwargs and state_dict keys
  compat  -- installs the `mmdet3d.ops` / `mmdet3d.models` import names when mmdet3d is absent

All compute runs in hand-written CUDA kernels of ``libpcreid_sm100.so`` (C ABI in ``include/pcreid.h``);
importing a compute entry point without the built library raises -- there is no CPU fallback.
"""
__version__ = "0.1.0"
