"""torch.library operators (namespace ``psg``) over the C ABI of libpsg_b200.so.

Each operator is registered for the CUDA dispatch key only: calling one with CPU tensors raises
(NotImplementedError from the dispatcher) -- there is no CPU or pure-PyTorch fallback.  The
operators allocate their outputs with the torch allocator, pass raw device pointers and the
current CUDA stream to one ``extern "C"`` entry point, and never synchronise.

Reference functions replaced (PointNet/models/pointnet_util.py): square_distance :19-40,
index_points :43-60, farthest_point_sample :63-84, query_ball_point :87-107, the 3-NN +
inverse-distance weights of PointNetFeaturePropagation.forward :301-307.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

_LIB = torch.library.Library("psg", "DEF")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"psg ops are float32 only (as the reference path is), got {t.dtype}")
    return t.contiguous()


# ---- farthest point sampling -----------------------------------------------------------------------
_LIB.define("fps(Tensor xyz, int npoint, Tensor start) -> Tensor")


def _fps_cuda(xyz, npoint, start):
    xyz = _f32c(xyz)
    B, N, three = xyz.shape
    assert three == 3
    st = start.to(device=xyz.device, dtype=torch.int32).contiguous()
    out = torch.empty(B, npoint, dtype=torch.int32, device=xyz.device)
    wsb = L.psg_fps_workspace(B, N)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=xyz.device)
    L.psg_fps(xyz.data_ptr(), B, B, N, npoint, st.data_ptr(), out.data_ptr(), None, ws.data_ptr(), wsb, _stream())
    return out.to(torch.int64)


_LIB.impl("fps", L.on_device_of(_fps_cuda), "CUDA")

# ---- square_distance -------------------------------------------------------------------------------
_LIB.define("square_distance(Tensor src, Tensor dst) -> Tensor")


def _sqd_cuda(src, dst):
    src, dst = _f32c(src), _f32c(dst)
    B, N, _ = src.shape
    M = dst.shape[1]
    out = torch.empty(B, N, M, dtype=torch.float32, device=src.device)
    L.psg_square_distance(src.data_ptr(), dst.data_ptr(), B, N, M, out.data_ptr(), _stream())
    return out


_LIB.impl("square_distance", L.on_device_of(_sqd_cuda), "CUDA")

# ---- ball query (one or two radii sharing a scan) --------------------------------------------------
_LIB.define("ball_query(float radius, int nsample, Tensor xyz, Tensor new_xyz) -> Tensor")
_LIB.define("ball_query2(float r0, int k0, float r1, int k1, Tensor xyz, Tensor new_xyz) -> (Tensor, Tensor)")


_GRID_MIN_N = 1024      # clouds at least this large go through the uniform-grid kernel (same results)


def _ball_launch(xyz, new_xyz, B, N, S, nr, r, k, o0, o1):
    if N >= _GRID_MIN_N:
        wsb = L.psg_ball_grid_workspace(B, N)
        ws = torch.empty(wsb, dtype=torch.uint8, device=xyz.device)
        L.psg_ball_query_grid(xyz.data_ptr(), B, B, N, new_xyz.data_ptr(), S, nr, r, k, o0.data_ptr(),
                              o1.data_ptr() if o1 is not None else None, ws.data_ptr(), wsb, _stream())
    else:
        L.psg_ball_query(xyz.data_ptr(), B, B, N, new_xyz.data_ptr(), S, nr, r, k, o0.data_ptr(),
                         o1.data_ptr() if o1 is not None else None, _stream())


def _ball_cuda(radius, nsample, xyz, new_xyz):
    xyz, new_xyz = _f32c(xyz), _f32c(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    out = torch.empty(B, S, nsample, dtype=torch.int32, device=xyz.device)
    r = (C.c_double * 2)(radius, 0.0)
    k = (C.c_int * 2)(nsample, 0)
    # the radius travels as a double: pointnet_util.py:102 squares the Python float in double
    # precision and only the comparison casts it to float32
    _ball_launch(xyz, new_xyz, B, N, S, 1, r, k, out, None)
    return out.to(torch.int64)


def _ball2_cuda(r0, k0, r1, k1, xyz, new_xyz):
    xyz, new_xyz = _f32c(xyz), _f32c(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    o0 = torch.empty(B, S, k0, dtype=torch.int32, device=xyz.device)
    o1 = torch.empty(B, S, k1, dtype=torch.int32, device=xyz.device)
    r = (C.c_double * 2)(r0, r1)
    k = (C.c_int * 2)(k0, k1)
    _ball_launch(xyz, new_xyz, B, N, S, 2, r, k, o0, o1)
    return o0.to(torch.int64), o1.to(torch.int64)


_LIB.impl("ball_query", L.on_device_of(_ball_cuda), "CUDA")
_LIB.impl("ball_query2", L.on_device_of(_ball2_cuda), "CUDA")

# ---- 3-NN + inverse-distance weights ---------------------------------------------------------------
_LIB.define("three_nn(Tensor xyz1, Tensor xyz2) -> (Tensor, Tensor, Tensor)")


def _three_nn_cuda(xyz1, xyz2):
    xyz1, xyz2 = _f32c(xyz1), _f32c(xyz2)
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    idx = torch.empty(B, N, 3, dtype=torch.int32, device=xyz1.device)
    w = torch.empty(B, N, 3, dtype=torch.float32, device=xyz1.device)
    d2 = torch.empty(B, N, 3, dtype=torch.float32, device=xyz1.device)
    L.psg_three_nn(xyz1.data_ptr(), B, B, N, xyz2.data_ptr(), S, idx.data_ptr(), w.data_ptr(), d2.data_ptr(), _stream())
    return idx.to(torch.int64), d2, w


_LIB.impl("three_nn", L.on_device_of(_three_nn_cuda), "CUDA")

# ---- index_points ----------------------------------------------------------------------------------
_LIB.define("index_points(Tensor points, Tensor idx) -> Tensor")


def _index_points_cuda(points, idx):
    points = _f32c(points)
    B, N, Cc = points.shape
    idx64 = idx.to(torch.int64).contiguous()
    M = idx64[0].numel()
    out = torch.empty(*idx64.shape, Cc, dtype=torch.float32, device=points.device)
    L.psg_index_points(points.data_ptr(), idx64.data_ptr(), B, N, Cc, M, out.data_ptr(), _stream())
    return out


_LIB.impl("index_points", L.on_device_of(_index_points_cuda), "CUDA")
