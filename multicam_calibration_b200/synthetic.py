"""Deterministic synthetic multi-camera calibration scenes.

Data generator for the tests and ``bench.py``; host-side numpy, not part of
the solve.  C cameras on a ring of radius 600 at height 200 looking at the
origin, a 5x7 board with 12.5 mm squares (same layout as the reference's
``generate_chessboard_objpoints``, detection.py:492-518), board poses
``rotvec ~ N(0, 0.6)``, ``trans ~ N(0, 60)``, pixel noise ``sigma``, missing
detections per (camera, frame) and/or per corner, and an initial guess that is
the truth perturbed.
"""
from dataclasses import dataclass

import numpy as np


def chessboard_objpoints(shape=(5, 7), square=12.5):
    """(rows*cols, 3) float64 board corners, z = 0 (detection.py:492-518 layout)."""
    rows, cols = shape
    pts = np.zeros((rows * cols, 3))
    pts[:, :2] = np.mgrid[0:rows, 0:cols].T.reshape(-1, 2) * square
    return pts


def _rotmat(r):
    r = np.asarray(r, dtype=float)
    th = np.linalg.norm(r, axis=-1)[..., None, None]
    k = r / np.where(th[..., 0] == 0, 1.0, th[..., 0])
    K = np.zeros(r.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
    K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def _rotvec(R):
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    th = np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1))
    n = np.linalg.norm(w)
    return w * th / (n if n > 0 else 1.0)


def forward_model(cam_params, poses, objpoints, chunk=4096):
    """Noise-free (C,F,N,2) projections for ``cam_params (C,12)``, ``poses (F,6)``."""
    C, F, N = len(cam_params), len(poses), len(objpoints)
    out = np.empty((C, F, N, 2))
    Rc = _rotmat(cam_params[:, 6:9])
    for s in range(0, F, chunk):
        p = poses[s:s + chunk]
        Xw = np.einsum("fij,nj->fni", _rotmat(p[:, :3]), objpoints) + p[:, None, 3:]
        for c in range(C):
            fx, fy, cx, cy, k1, k2 = cam_params[c, :6]
            Xc = Xw @ Rc[c].T + cam_params[c, 9:12]
            x, y = Xc[..., 0] / Xc[..., 2], Xc[..., 1] / Xc[..., 2]
            r2 = x * x + y * y
            d = 1 + k1 * r2 + k2 * r2 * r2
            out[c, s:s + chunk, :, 0] = fx * x * d + cx
            out[c, s:s + chunk, :, 1] = fy * y * d + cy
    return out


@dataclass
class Scene:
    uvs: np.ndarray            # (C,F,N,2) float64, NaN = missing
    objpoints: np.ndarray      # (N,3)
    true_cams: np.ndarray      # (C,12)
    true_poses: np.ndarray     # (F,6)
    init_cams: np.ndarray      # (C,12)
    init_poses: np.ndarray     # (F,6)

    @staticmethod
    def _split(cams):
        intr = []
        for p in cams:
            K = np.eye(3)
            K[0, 0], K[1, 1], K[0, 2], K[1, 2] = p[:4]
            intr.append((K, np.r_[p[4:6], 0.0, 0.0, 0.0]))
        return cams[:, 6:].copy(), intr

    def init_args(self):
        """Positional arguments of ``bundle_adjust`` (reference signature)."""
        ext, intr = self._split(self.init_cams)
        return self.uvs, ext, intr, self.objpoints, self.init_poses

    def x0(self):
        return np.concatenate([self.init_cams.ravel(), self.init_poses.ravel()])

    def x_true(self):
        return np.concatenate([self.true_cams.ravel(), self.true_poses.ravel()])

    @property
    def n_obs(self):
        return int((~np.isnan(self.uvs).any(-1)).sum())


def make_scene(n_cameras=6, n_frames=500, board=(5, 7), square=12.5, sigma=0.3,
               p_missing_view=0.0, p_missing_corner=0.0, seed=0, perturb=1.0, shard=0):
    """``seed`` fixes the cameras; ``(seed, shard)`` fixes the frames, so ranks that
    generate different shards of one scene share the cameras."""
    rng = np.random.default_rng(seed)
    obj = chessboard_objpoints(board, square)
    C = n_cameras
    cams = np.zeros((C, 12))
    for c in range(C):
        phi = 2 * np.pi * c / C
        pos = np.array([600 * np.cos(phi), 600 * np.sin(phi), 200.0])
        z = -pos / np.linalg.norm(pos)
        xax = np.cross(z, [0.0, 0.0, 1.0])
        xax /= np.linalg.norm(xax)
        yax = np.cross(z, xax)
        R = np.stack([xax, yax, z])            # rows: camera axes in world coords
        cams[c, 6:9] = _rotvec(R)
        cams[c, 9:12] = -R @ pos
    cams[:, 0] = rng.normal(1200, 20, C)
    cams[:, 1] = rng.normal(1200, 20, C)
    cams[:, 2] = rng.normal(640, 5, C)
    cams[:, 3] = rng.normal(512, 5, C)
    cams[:, 4] = rng.normal(-0.1, 0.02, C)
    cams[:, 5] = rng.normal(0.05, 0.01, C)
    init_cams = cams.copy()
    init_cams[:, 6:9] += perturb * rng.normal(0, 0.01, (C, 3))
    init_cams[:, 9:12] += perturb * rng.normal(0, 3, (C, 3))
    init_cams[:, 0:2] += perturb * rng.uniform(-10, 10, (C, 2))
    init_cams[:, 4:6] *= 1 - 0.5 * min(perturb, 1.0)
    rng = np.random.default_rng([seed, 1, shard])
    poses = np.concatenate([rng.normal(0, 0.6, (n_frames, 3)),
                            rng.normal(0, 60, (n_frames, 3))], axis=1)
    uvs = forward_model(cams, poses, obj)
    if sigma:
        uvs += rng.normal(0, sigma, uvs.shape)
    if p_missing_view:
        uvs[rng.random((C, n_frames)) < p_missing_view] = np.nan
    if p_missing_corner:
        uvs[rng.random(uvs.shape[:3]) < p_missing_corner] = np.nan
    init_poses = poses.copy()
    init_poses[:, :3] += perturb * rng.normal(0, 0.02, (n_frames, 3))
    init_poses[:, 3:] += perturb * rng.normal(0, 2, (n_frames, 3))
    return Scene(uvs, obj, cams, poses, init_cams, init_poses)


def make_keypoints(n_points=1_000_000, n_cameras=6, sigma=0.3, p_missing=0.2, seed=0):
    """Config 5: world points ~ N(0, 80) seen by the ring cameras, per-view NaNs.

    Returns ``(all_uvs list of (P,2), extrinsics (C,6), intrinsics list, points (P,3))``.
    """
    sc = make_scene(n_cameras, 1, sigma=0.0, seed=seed)
    rng = np.random.default_rng(seed + 1)
    pts = rng.normal(0, 80, (n_points, 3))
    ext, intr = Scene._split(sc.true_cams)
    all_uvs = []
    Rc = _rotmat(sc.true_cams[:, 6:9])
    for c in range(n_cameras):
        fx, fy, cx, cy, k1, k2 = sc.true_cams[c, :6]
        Xc = pts @ Rc[c].T + sc.true_cams[c, 9:12]
        x, y = Xc[:, 0] / Xc[:, 2], Xc[:, 1] / Xc[:, 2]
        r2 = x * x + y * y
        d = 1 + k1 * r2 + k2 * r2 * r2
        uv = np.stack([fx * x * d + cx, fy * y * d + cy], axis=-1)
        uv += rng.normal(0, sigma, uv.shape)
        uv[rng.random(n_points) < p_missing] = np.nan
        all_uvs.append(uv)
    return all_uvs, ext, intr, pts


def make_camera_poses(n_cameras=6, n_frames=500, p_missing=0.25, sigma_rot=0.01, sigma_trans=1.0, seed=0):
    """Per-camera board poses as ``estimate_pose`` (cv2.solvePnP per camera and frame,
    calibration.py:72-113) would return them: ``T_board->cam = T_world->cam T_board->world`` of a
    :func:`make_scene` rig, perturbed by a small rigid noise, NaN rows where the camera did not
    see the complete board.  Returns ``(all_calib_poses (C,F,6), true extrinsics (C,6), true
    board poses (F,6))`` with camera 0 as the world frame (what estimate_all_extrinsics recovers)."""
    rng = np.random.default_rng([seed, 7])
    cams = make_scene(n_cameras, 1, sigma=0.0, seed=seed).true_cams
    board_world = np.concatenate([rng.normal(0, 0.6, (n_frames, 3)), rng.normal(0, 60, (n_frames, 3))], axis=1)

    def mat(v):
        T = np.zeros(v.shape[:-1] + (4, 4))
        T[..., :3, :3] = _rotmat(v[..., :3])
        T[..., :3, 3] = v[..., 3:]
        T[..., 3, 3] = 1.0
        return T

    def vec(T):
        R = T[..., :3, :3]
        w = np.stack([R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]], axis=-1)
        th = np.arccos(np.clip((np.trace(R, axis1=-2, axis2=-1) - 1) / 2, -1, 1))[..., None]
        n = np.linalg.norm(w, axis=-1, keepdims=True)
        return np.concatenate([w * th / np.where(n > 0, n, 1.0), T[..., :3, 3]], axis=-1)
    T_wc = mat(cams[:, 6:12])                         # world -> camera
    T_bw = mat(board_world)                           # board -> world
    T_root = np.linalg.inv(T_wc[0])                   # camera 0 becomes the world frame
    ext = vec(T_wc @ T_root)
    board = vec(np.linalg.inv(T_root) @ T_bw)
    noise = np.concatenate([rng.normal(0, sigma_rot, (n_cameras, n_frames, 3)),
                            rng.normal(0, sigma_trans, (n_cameras, n_frames, 3))], axis=-1)
    poses = vec(mat(noise) @ (T_wc[:, None] @ T_bw[None]))
    poses[rng.random((n_cameras, n_frames)) < p_missing] = np.nan
    return poses, ext, board
