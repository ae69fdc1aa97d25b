"""Drop-in for ``multicam_calibration.geometry`` (reference geometry.py:8-433).

Same names, argument order, broadcasting and NaN conventions.  The batched
functions of the hot path (``project_points``, ``undistort_points``,
``triangulate``) run on the GPU through ``libmcba``; the small algebraic
helpers (a handful of 3x3 / 4x4 operations the reference calls once per camera
or frame) are host numpy, as SURVEY.md section 2 scopes them.
"""
import ctypes

import numpy as np

from . import _native
from ._native import check

na = np.newaxis


# ----------------------------------------------------------------------------
# small host helpers (API surface)
# ----------------------------------------------------------------------------
def rodrigues(r):
    """Rotation vector (...,3) -> matrix (...,3,3) (geometry.py:8-35)."""
    r = np.asarray(r, dtype=float)
    theta = np.sqrt((r * r).sum(-1))
    k = r / np.where(theta == 0, 1.0, theta)[..., na]
    kx, ky, kz = k[..., 0], k[..., 1], k[..., 2]
    zero = np.zeros_like(kx)
    Kmat = np.stack([np.stack([zero, -kz, ky], -1),
                     np.stack([kz, zero, -kx], -1),
                     np.stack([-ky, kx, zero], -1)], -2)
    s = np.sin(theta)[..., na, na]
    c = (1 - np.cos(theta))[..., na, na]
    return s * Kmat + c * (Kmat @ Kmat) + np.eye(3)


def rodrigues_inv(R):
    """Rotation matrix (...,3,3) -> vector (...,3) (geometry.py:38-65)."""
    R = np.asarray(R, dtype=float)
    axis = np.stack([R[..., 2, 1] - R[..., 1, 2],
                     R[..., 0, 2] - R[..., 2, 0],
                     R[..., 1, 0] - R[..., 0, 1]], axis=-1)
    angle = np.arccos((np.trace(R, axis1=-2, axis2=-1) - 1) / 2)[..., na]
    norm = np.linalg.norm(axis, axis=-1, keepdims=True)
    norm = norm + (norm == 0)
    return axis * angle / norm


def get_transformation_matrix(t):
    """(...,6) -> (...,4,4) (geometry.py:155-175)."""
    t = np.asarray(t, dtype=float)
    T = np.zeros((*t.shape[:-1], 4, 4))
    T[..., :3, :3] = rodrigues(t[..., :3])
    T[..., :3, 3] = t[..., 3:]
    T[..., 3, 3] = 1
    return T


def get_transformation_vector(T):
    """(...,4,4) -> (...,6) (geometry.py:178-197)."""
    T = np.asarray(T, dtype=float)
    return np.concatenate([rodrigues_inv(T[..., :3, :3]), T[..., :3, 3]], axis=-1)


def euclidean_to_homogenous(x_euclidean):
    """Append a one (geometry.py:232-252)."""
    x = np.asarray(x_euclidean)
    return np.concatenate((x, np.ones((*x.shape[:-1], 1))), axis=-1)


def homogeneous_to_euclidean(x_homogenous):
    """Divide by the last coordinate (geometry.py:255-274)."""
    x = np.asarray(x_homogenous)
    return x[..., :-1] / x[..., -1:]


def apply_rigid_transform(transform, points):
    """Transform (6,) or (...,4,4) applied to points (...,3) (geometry.py:128-152)."""
    transform = np.asarray(transform, dtype=float)
    if transform.shape == (6,):
        transform = get_transformation_matrix(transform)
    hom = euclidean_to_homogenous(points)
    return np.matmul(transform, hom[..., na])[..., :3, 0]


def get_projection_matrix(extrinsics, intrinsics):
    """``P = K [R | t]`` (geometry.py:200-229)."""
    camera_matrix, _ = intrinsics
    return np.matmul(camera_matrix, get_transformation_matrix(np.asarray(extrinsics, dtype=float))[:3])


def rigid_transform_from_correspondences(source_points, target_points):
    """Kabsch fit, returns (6-vector, rmsd) (geometry.py:68-125)."""
    src = np.asarray(source_points, dtype=float).reshape(-1, 3)
    dst = np.asarray(target_points, dtype=float).reshape(-1, 3)
    mu_s, mu_d = src.mean(0), dst.mean(0)
    U, _, Vt = np.linalg.svd((src - mu_s).T @ (dst - mu_d))
    R = Vt.T @ U.T
    if np.linalg.det(R) < 0:
        Vt[-1, :] *= -1
        R = Vt.T @ U.T
    t = np.concatenate((rodrigues_inv(R), mu_d - R @ mu_s))
    moved = apply_rigid_transform(t, src)
    rmsd = np.sqrt(np.mean(np.sum((moved - dst) ** 2, axis=1)))
    return t, rmsd


# ----------------------------------------------------------------------------
# GPU-backed batched functions
# ----------------------------------------------------------------------------
def _device_ctx():
    torch = _native.require_cuda()
    return torch, _native.load(), torch.cuda.current_device()


def _h(a):
    return np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(ctypes.c_void_p)


def project_points(points, extrinsics, camera_matrix, dist_coefs=None):
    """Project (...,3) world points to (...,2) pixels (geometry.py:277-325)."""
    torch, lib, dev = _device_ctx()
    pts = np.ascontiguousarray(points, dtype=np.float64)
    lead = pts.shape[:-1]
    d_pts = _native.to_device(pts.reshape(-1, 3), dev)
    d_uv = torch.empty((d_pts.shape[0], 2), dtype=torch.float64, device=d_pts.device)
    ext = np.ascontiguousarray(extrinsics, dtype=np.float64)
    K = np.ascontiguousarray(camera_matrix, dtype=np.float64)
    dist = None if dist_coefs is None else np.ascontiguousarray(np.asarray(dist_coefs, dtype=np.float64)[:2])
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    check(lib.mcba_project_points(dev, stream, ctypes.c_void_p(d_pts.data_ptr()), d_pts.shape[0],
                                  _h(ext), _h(K), None if dist is None else _h(dist),
                                  ctypes.c_void_p(d_uv.data_ptr())))
    return _native.to_host(d_uv).reshape(*lead, 2)


def project_points_multi(points, all_extrinsics, all_intrinsics, _device_points=None):
    """``np.stack([project_points(points, ext, K, dist) for ...])`` in one pass over the points
    (``mcba_project_points_multi``): (...,3) -> (C,...,2).  ``all_intrinsics`` is the reference's
    list of ``(camera_matrix, dist_coefs)``; every camera must either have distortion or not."""
    torch, lib, dev = _device_ctx()
    if _device_points is None:
        pts = np.ascontiguousarray(points, dtype=np.float64)
        lead = pts.shape[:-1]
        d_pts = _native.to_device(pts.reshape(-1, 3), dev)
    else:
        d_pts, lead = _device_points
    C = len(all_extrinsics)
    ext = np.ascontiguousarray(np.stack([np.asarray(e, dtype=np.float64) for e in all_extrinsics]))
    Ks = np.ascontiguousarray(np.stack([np.asarray(K, dtype=np.float64) for K, _ in all_intrinsics]))
    has = [d is not None for _, d in all_intrinsics]
    if any(has) != all(has):
        raise ValueError("project_points_multi: dist_coefs must be given for every camera or for none")
    dist = None
    if all(has):
        dist = np.ascontiguousarray(np.stack([np.asarray(d, dtype=np.float64).ravel()[:2] for _, d in all_intrinsics]))
    d_uv = torch.empty((C, d_pts.shape[0], 2), dtype=torch.float64, device=d_pts.device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    check(lib.mcba_project_points_multi(dev, stream, ctypes.c_void_p(d_pts.data_ptr()), d_pts.shape[0], C,
                                        _h(ext), _h(Ks), None if dist is None else _h(dist),
                                        ctypes.c_void_p(d_uv.data_ptr())))
    return _native.to_host(d_uv).reshape(C, *lead, 2)


def undistort_points(uvs, camera_matrix, dist_coefs):
    """NaN-aware ``cv2.undistortPoints(uv, K, dist, None, K)`` (geometry.py:328-358)."""
    torch, lib, dev = _device_ctx()
    uv = np.ascontiguousarray(uvs, dtype=np.float64)
    shape = uv.shape
    d_in = _native.to_device(uv.reshape(-1, 2), dev)
    d_out = torch.empty_like(d_in)
    dist = np.zeros(5)
    dc = np.asarray(dist_coefs, dtype=np.float64).ravel()[:5]
    dist[:dc.size] = dc
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    check(lib.mcba_undistort_points(dev, stream, ctypes.c_void_p(d_in.data_ptr()), d_in.shape[0],
                                    _h(camera_matrix), _h(dist), ctypes.c_void_p(d_out.data_ptr())))
    return _native.to_host(d_out).reshape(shape)


def triangulate(all_uvs, all_extrinsics, all_intrinsics):
    """Robust all-pairs triangulation (geometry.py:361-433): list of C arrays
    (P,2) with NaN = missing -> (P,3), NaN where fewer than two views."""
    torch, lib, dev = _device_ctx()
    uv = np.ascontiguousarray(np.stack([np.asarray(u, dtype=np.float64) for u in all_uvs]))
    C, P, _ = uv.shape
    ext = np.ascontiguousarray(np.stack([np.asarray(e, dtype=np.float64) for e in all_extrinsics]))
    Ks = np.ascontiguousarray(np.stack([np.asarray(K, dtype=np.float64) for K, _ in all_intrinsics]))
    dists = np.zeros((C, 5))
    for c, (_, dc) in enumerate(all_intrinsics):
        dc = np.asarray(dc, dtype=np.float64).ravel()[:5]
        dists[c, :dc.size] = dc
    d_uv = _native.to_device(uv, dev)
    d_out = torch.empty((P, 3), dtype=torch.float64, device=d_uv.device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    check(lib.mcba_triangulate(dev, stream, ctypes.c_void_p(d_uv.data_ptr()), C, P, _h(ext), _h(Ks),
                               _h(dists), ctypes.c_void_p(d_out.data_ptr())))
    return _native.to_host(d_out)
