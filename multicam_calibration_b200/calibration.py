"""Initialisation algebra of ``multicam_calibration.calibration`` on the B200 engine
(reference calibration.py:116-277): the rigid-transform
step that turns per-camera board poses into the extrinsics and consensus board poses
``bundle_adjust`` starts from.

Same names, argument order and return values as the reference.  The per-frame work
(relative transforms, world-frame board poses, the medians over frames / cameras) runs
in ``libmcba`` (``csrc/k6_init.cu``); the camera graph (a handful of nodes) stays on the
host.  OpenCV-bound steps of the reference (``get_intrinsics``, ``estimate_pose``:
``cv2.calibrateCamera`` / ``cv2.solvePnP``) are out of scope and not provided.
"""
import ctypes

import numpy as np

from . import _native
from ._native import check
from .geometry import get_transformation_matrix, get_transformation_vector


def _ctx():
    torch = _native.require_cuda()
    dev = torch.cuda.current_device()
    return torch, _native.load(), dev, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _pairwise_device(lib, dev, stream, d_poses1, d_poses2, n_frames):
    out = (ctypes.c_double * 6)()
    n_common = ctypes.c_int64()
    check(lib.mcba_pairwise_transform(dev, stream, ctypes.c_void_p(d_poses1.data_ptr()),
                                      ctypes.c_void_p(d_poses2.data_ptr()), n_frames, out, ctypes.byref(n_common)))
    return np.array(out[:], dtype=np.float64)


def estimate_pairwise_camera_transform(camera1_poses, camera2_poses):
    """Median over the common frames of ``T2 T1^-1`` in vector form: the transform from camera
    1's to camera 2's coordinates (calibration.py:116-143)."""
    torch, lib, dev, stream = _ctx()
    p1 = np.ascontiguousarray(camera1_poses, dtype=np.float64)
    p2 = np.ascontiguousarray(camera2_poses, dtype=np.float64)
    if p1.ndim != 2 or p1.shape[1] != 6 or p1.shape != p2.shape:
        raise ValueError("camera poses must both be (n_frames, 6)")
    return _pairwise_device(lib, dev, stream, _native.to_device(p1, dev), _native.to_device(p2, dev), p1.shape[0])


def get_camera_spanning_tree(all_calib_poses, root=0):
    """Maximal spanning tree of the camera graph weighted by co-detected frames, as edges
    ``(nearer to root, farther)`` ordered by distance from the root (calibration.py:146-197).

    The reference builds it with networkx (Kruskal on the edges ``(i, j), i < j`` sorted by
    decreasing weight, ties in insertion order; ``Graph.edges`` iteration order; a stable sort by
    the root distance of the first node); this is the same procedure on plain lists."""
    poses = np.asarray(all_calib_poses)
    C = len(poses)
    detected = ~np.isnan(poses).any(2)
    common = detected.astype(np.int64) @ detected.astype(np.int64).T
    edges = [(int(common[i, j]), i, j) for i in range(C) for j in range(i + 1, C)]
    edges.sort(key=lambda e: e[0], reverse=True)      # stable: ties keep (i, j) order
    parent = list(range(C))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    adjacency = {c: [] for c in range(C)}             # insertion-ordered, like networkx's adjacency dicts
    for _, u, v in edges:
        ru, rv = find(u), find(v)
        if ru != rv:
            parent[ru] = rv
            adjacency[u].append(v)
            adjacency[v].append(u)
    # hop distance from the root (the tree is connected: the camera graph is complete)
    dist = {root: 0}
    frontier = [root]
    while frontier:
        nxt = []
        for u in frontier:
            for v in adjacency[u]:
                if v not in dist:
                    dist[v] = dist[u] + 1
                    nxt.append(v)
        frontier = nxt
    seen, tree = set(), []
    for u in range(C):                                # Graph.edges: each edge once, from its first-listed node
        for v in adjacency[u]:
            if v not in seen:
                tree.append(tuple(sorted((u, v), key=lambda n: dist[n])))
        seen.add(u)
    tree.sort(key=lambda e: dist[e[0]])
    return tree


def estimate_all_extrinsics(all_calib_poses, root=0):
    """Transforms from the root camera to every camera, chained along the spanning tree
    (calibration.py:200-242).  Returns ``(all_extrinsics (C,6), spanning_tree)``."""
    torch, lib, dev, stream = _ctx()
    poses = np.ascontiguousarray(all_calib_poses, dtype=np.float64)
    if poses.ndim != 3 or poses.shape[2] != 6:
        raise ValueError("all_calib_poses must be (n_cameras, n_frames, 6)")
    C, F, _ = poses.shape
    d_poses = _native.to_device(poses, dev)
    spanning_tree = get_camera_spanning_tree(poses, root=root)
    matrices = [None] * C
    matrices[root] = np.eye(4)
    for c1, c2 in spanning_tree:
        transform = _pairwise_device(lib, dev, stream, d_poses[c1], d_poses[c2], F)
        matrices[c2] = get_transformation_matrix(transform) @ matrices[c1]
    return np.array([get_transformation_vector(T) for T in matrices]), spanning_tree


def consensus_calib_poses(all_calib_poses, all_extrinsics):
    """Per frame, the median over the detecting cameras of the board pose mapped to world
    coordinates (calibration.py:245-277); NaN rows where no camera saw the board."""
    torch, lib, dev, stream = _ctx()
    poses = np.ascontiguousarray(all_calib_poses, dtype=np.float64)
    ext = np.ascontiguousarray(all_extrinsics, dtype=np.float64)
    if poses.ndim != 3 or poses.shape[2] != 6 or ext.shape != (poses.shape[0], 6):
        raise ValueError("all_calib_poses must be (n_cameras, n_frames, 6) and all_extrinsics (n_cameras, 6)")
    C, F, _ = poses.shape
    d_poses = _native.to_device(poses, dev)
    d_out = torch.empty((F, 6), dtype=torch.float64, device=d_poses.device)
    check(lib.mcba_consensus_poses(dev, stream, ctypes.c_void_p(d_poses.data_ptr()),
                                   ext.ctypes.data_as(ctypes.c_void_p), C, F, ctypes.c_void_p(d_out.data_ptr())))
    return _native.to_host(d_out)
