"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/libops_oracle.so (CPU restatement of the five
mmdet3d point ops, see ops_oracle.c) and of oracle/_ref/libref_ops.so (the reference's own .cu files
compiled unmodified for sm_100a; GPU box only).  Signatures mirror the reference python wrappers:
  furthest_point_sample.py:7-78, knn.py:7-71, ball_query.py:7-54, group_points.py:169-220,
  gather_points.py:7-50  (all under mmdet3d/ops/).
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_CPU = None
_REF = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def _cpu():
    global _CPU
    if _CPU is None:
        path = os.path.join(_HERE, "libops_oracle.so")
        if not os.path.exists(path):
            build()
        _CPU = ctypes.CDLL(path)
    return _CPU


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_ops.so")) and torch.cuda.is_available()


def _ref():
    global _REF
    if _REF is None:
        _REF = ctypes.CDLL(os.path.join(_HERE, "_ref", "libref_ops.so"))
    return _REF


def _fp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(t, dtype):
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(dtype, copy=False))


# ---------------------------------------------------------------- CPU restatement
def fps_block_size(n):
    return _cpu().oracle_fps_block_size(int(n))


def furthest_point_sample(xyz, m):
    x = _c(xyz, np.float32)
    B, N, _ = x.shape
    temp = np.full((B, N), 1e10, np.float32)
    idx = np.zeros((B, m), np.int32)
    _cpu().oracle_fps(B, N, m, _fp(x), _fp(temp), _fp(idx))
    return torch.from_numpy(idx)


def furthest_point_sample_with_dist(dist, m):
    x = _c(dist, np.float32)
    B, N, _ = x.shape
    temp = np.full((B, N), 1e10, np.float32)
    idx = np.zeros((B, m), np.int32)
    _cpu().oracle_fps_with_dist(B, N, m, _fp(x), _fp(temp), _fp(idx))
    return torch.from_numpy(idx)


def pairwise_sqdist(a, b, norm=False):
    """calc_square_dist (furthest_point_sample/utils.py:4-31) in the kernel's arithmetic: a (B,N,C), b (B,M,C) -> (B,N,M)."""
    x, y = _c(a, np.float32), _c(b, np.float32)
    B, N, C = x.shape
    M = y.shape[1]
    out = np.zeros((B, N, M), np.float32)
    _cpu().oracle_pairwise_sqdist(B, N, M, C, _fp(x), _fp(y), _fp(out), int(bool(norm)))
    return torch.from_numpy(out)


def knn(k, xyz, center_xyz=None, transposed=False, return_dist=False):
    """-> int32 [B,k,S] (knn.py:62 transposes the kernel's [B,S,k])."""
    if center_xyz is None:
        center_xyz = xyz
    if transposed:
        xyz = xyz.transpose(2, 1)
        center_xyz = center_xyz.transpose(2, 1)
    x = _c(xyz, np.float32)
    c = _c(center_xyz, np.float32)
    B, N, _ = x.shape
    S = c.shape[1]
    idx = np.zeros((B, S, k), np.int32)
    d2 = np.zeros((B, S, k), np.float32)
    _cpu().oracle_knn(B, N, S, k, _fp(x), _fp(c), _fp(idx), _fp(d2))
    out = torch.from_numpy(idx).transpose(2, 1).contiguous()
    if return_dist:
        return out, torch.from_numpy(d2)
    return out


def ball_query(min_radius, max_radius, sample_num, xyz, center_xyz):
    x = _c(xyz, np.float32)
    c = _c(center_xyz, np.float32)
    B, N, _ = x.shape
    S = c.shape[1]
    idx = np.zeros((B, S, sample_num), np.int32)
    f = _cpu().oracle_ball_query
    f.argtypes = [ctypes.c_int] * 3 + [ctypes.c_float] * 2 + [ctypes.c_int] + [ctypes.c_void_p] * 3
    f(B, N, S, float(min_radius), float(max_radius), sample_num, _fp(c), _fp(x), _fp(idx))
    return torch.from_numpy(idx)


def grouping_operation(features, indices):
    f = _c(features, np.float32)
    i = _c(indices, np.int32)
    B, C, N = f.shape
    _, S, K = i.shape
    out = np.zeros((B, C, S, K), np.float32)
    _cpu().oracle_group_points(B, C, N, S, K, _fp(f), _fp(i), _fp(out))
    return torch.from_numpy(out)


def gather_points(features, indices):
    f = _c(features, np.float32)
    i = _c(indices, np.int32)
    B, C, N = f.shape
    M = i.shape[1]
    out = np.zeros((B, C, M), np.float32)
    _cpu().oracle_gather_points(B, C, N, M, _fp(f), _fp(i), _fp(out))
    return torch.from_numpy(out)


def three_nn_dist2(target, source):
    """-> (dist2 (B,N,3) squared distances, idx int32 (B,N,3)): what the reference kernel writes."""
    t, s = _c(target, np.float32), _c(source, np.float32)
    B, N, _ = t.shape
    M = s.shape[1]
    d2 = np.zeros((B, N, 3), np.float32)
    idx = np.zeros((B, N, 3), np.int32)
    _cpu().oracle_three_nn(B, N, M, _fp(t), _fp(s), _fp(d2), _fp(idx))
    return torch.from_numpy(d2), torch.from_numpy(idx)


def three_nn(target, source):
    """-> (dist (B,N,3) = sqrt(dist2), idx int32 (B,N,3)) as mmdet3d/ops/interpolate/three_nn.py:9-46."""
    d2, idx = three_nn_dist2(target, source)
    return torch.sqrt(d2), idx


def three_interpolate(features, indices, weight):
    f, i, w = _c(features, np.float32), _c(indices, np.int32), _c(weight, np.float32)
    B, C, M = f.shape
    N = i.shape[1]
    out = np.zeros((B, C, N), np.float32)
    _cpu().oracle_three_interpolate(B, C, M, N, _fp(f), _fp(i), _fp(w), _fp(out))
    return torch.from_numpy(out)


# ---------------------------------------------------------------- reference .cu on the GPU (oracle/_ref)
def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ref_furthest_point_sample(xyz, m):
    B, N, _ = xyz.shape
    temp = torch.full((B, N), 1e10, device=xyz.device)
    idx = torch.zeros(B, m, dtype=torch.int32, device=xyz.device)
    _ref().ref_fps(B, N, m, _p(xyz), _p(temp), _p(idx), _s())
    return idx


def ref_furthest_point_sample_with_dist(dist, m):
    B, N, _ = dist.shape
    temp = torch.full((B, N), 1e10, device=dist.device)
    idx = torch.zeros(B, m, dtype=torch.int32, device=dist.device)
    _ref().ref_fps_with_dist(B, N, m, _p(dist), _p(temp), _p(idx), _s())
    return idx


def ref_knn(k, xyz, center_xyz=None, return_dist=False):
    if center_xyz is None:
        center_xyz = xyz
    B, N, _ = xyz.shape
    S = center_xyz.shape[1]
    idx = torch.zeros(B, S, k, dtype=torch.int32, device=xyz.device)
    d2 = torch.zeros(B, S, k, device=xyz.device)
    _ref().ref_knn(B, N, S, k, _p(xyz), _p(center_xyz), _p(idx), _p(d2), _s())
    out = idx.transpose(2, 1).contiguous()
    return (out, d2) if return_dist else out


def ref_ball_query(min_radius, max_radius, sample_num, xyz, center_xyz):
    B, N, _ = xyz.shape
    S = center_xyz.shape[1]
    idx = torch.zeros(B, S, sample_num, dtype=torch.int32, device=xyz.device)
    f = _ref().ref_ball_query
    f.argtypes = [ctypes.c_int] * 3 + [ctypes.c_float] * 2 + [ctypes.c_int] + [ctypes.c_void_p] * 4
    f(B, N, S, float(min_radius), float(max_radius), sample_num, _p(center_xyz), _p(xyz), _p(idx), _s())
    return idx


def ref_grouping_operation(features, indices):
    B, C, N = features.shape
    _, S, K = indices.shape
    out = torch.zeros(B, C, S, K, device=features.device)
    _ref().ref_group_points(B, C, N, S, K, _p(features), _p(indices), _p(out), _s())
    return out


def ref_gather_points(features, indices):
    B, C, N = features.shape
    M = indices.shape[1]
    out = torch.zeros(B, C, M, device=features.device)
    _ref().ref_gather_points(B, C, N, M, _p(features), _p(indices), _p(out), _s())
    return out


def ref_three_nn(target, source):
    B, N, _ = target.shape
    M = source.shape[1]
    d2 = torch.zeros(B, N, 3, device=target.device)
    idx = torch.zeros(B, N, 3, dtype=torch.int32, device=target.device)
    _ref().ref_three_nn(B, N, M, _p(target), _p(source), _p(d2), _p(idx), _s())
    return torch.sqrt(d2), idx


def ref_three_interpolate(features, indices, weight):
    B, C, M = features.shape
    N = indices.shape[1]
    out = torch.zeros(B, C, N, device=features.device)
    _ref().ref_three_interpolate(B, C, M, N, _p(features), _p(indices), _p(weight), _p(out), _s())
    return out


# ---------------------------------------------------------------- canonical distance arithmetic (CPU)
def sqdist_expand(new_xyz, xyz):
    """C restatement of square_distance(new_xyz, xyz) -> (B, S, N); must equal torch bit for bit."""
    x = _c(xyz, np.float32)
    c = _c(new_xyz, np.float32)
    B, N, _ = x.shape
    S = c.shape[1]
    out = np.zeros((B, S, N), np.float32)
    _cpu().oracle_sqdist_expand(B, N, S, _fp(x), _fp(c), _fp(out))
    return torch.from_numpy(out)


def dgcnn_pd(x):
    """C restatement of dgcnn_orig.knn's pairwise_distance for x (B, C, N) -> (B, N, N)."""
    a = _c(x, np.float32)
    B, C, N = a.shape
    out = np.zeros((B, N, N), np.float32)
    _cpu().oracle_dgcnn_pd(B, C, N, _fp(a), _fp(out))
    return torch.from_numpy(out)
