"""Frame-sharded multi-GPU bundle adjustment.

One process per GPU (``torchrun``).  Frames are independent given the camera
parameters, so each rank owns a contiguous range of frames, runs the fused
residual/Jacobian/Schur kernels on its range, and only the packed reduced camera
system ``[S | b | g | diag | scalars]`` (5.4k doubles at 6 cameras) is summed
with one NCCL all-reduce per evaluation inside ``libmcba``.  Every rank solves
the small system redundantly (identical inputs, so no broadcast) and
back-substitutes its own poses.  ``torch.distributed`` provides the rendezvous:
the ncclUniqueId broadcast and the final gather of poses.
"""
import os

import numpy as np


def _dist():
    import torch.distributed as dist
    return dist


def world_size():
    dist = _dist()
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    dist = _dist()
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_bounds(n_frames, world, r):
    """Contiguous balanced partition: the first ``n_frames % world`` ranks get one extra frame."""
    if world < 1 or not 0 <= r < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(int(n_frames), world)
    start = r * base + min(r, extra)
    return start, start + base + (1 if r < extra else 0)


def split_params(x, n_cameras, start, stop):
    """Camera block (replicated) + this rank's pose blocks."""
    nc = 12 * n_cameras
    x = np.asarray(x, dtype=np.float64)
    return np.concatenate([x[:nc], x[nc + 6 * start: nc + 6 * stop]])


def merge_params(cam_block, pose_blocks):
    return np.concatenate([np.asarray(cam_block)] + [np.asarray(p).ravel() for p in pose_blocks])


def local_device():
    import torch
    lr = int(os.environ.get("LOCAL_RANK", rank()))
    return lr % max(torch.cuda.device_count(), 1)


def broadcast_unique_id():
    """ncclUniqueId made on rank 0 and shared through torch.distributed (any backend)."""
    import ctypes
    from . import _native
    dist = _dist()
    box = [None]
    if rank() == 0:
        buf = ctypes.create_string_buffer(128)
        _native.check(_native.load().mcba_comm_unique_id(buf))
        box[0] = bytes(buf.raw)
    dist.broadcast_object_list(box, src=0)
    return box[0]


def gather_arrays(local):
    """All ranks receive the list of every rank's array (host objects; poses are small)."""
    dist = _dist()
    out = [None] * world_size()
    dist.all_gather_object(out, np.asarray(local))
    return out


def solve_sharded(uvs_used, calib_objpoints, x0, **opt_kwargs):
    """Solve with frames sharded over the ranks of the default process group.

    ``uvs_used`` (C,F,N,2) and ``x0`` are the full (replicated) problem; the return
    value ``(x, result)`` is the full solution on every rank.
    """
    import torch
    from .engine import BAProblem
    W, r = world_size(), rank()
    C, F = uvs_used.shape[0], uvs_used.shape[1]
    if F < W:
        raise ValueError(f"{F} frames cannot be sharded over {W} ranks")
    start, stop = shard_bounds(F, W, r)
    dev = local_device()
    torch.cuda.set_device(dev)
    uid = broadcast_unique_id()
    prob = BAProblem(uvs_used[:, start:stop], calib_objpoints, device=dev, comm=(uid, r, W))
    try:
        x_loc, result = prob.solve(split_params(x0, C, start, stop), **opt_kwargs)
        result["peer_memory"] = bool(getattr(prob, "peer_memory", False))
    finally:
        prob.close()
    nc = 12 * C
    poses = gather_arrays(x_loc[nc:])
    grads = gather_arrays(result.grad[nc:])
    x = merge_params(x_loc[:nc], poses)
    result["x"] = x
    result["grad"] = merge_params(result.grad[:nc], grads)
    result["active_mask"] = np.zeros_like(x)
    result["fun"] = None   # the residual vector interleaves ranks per camera; not gathered
    result["shard"] = (start, stop)
    return x, result
