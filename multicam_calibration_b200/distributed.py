"""Frame-sharded multi-GPU bundle adjustment.

One process per GPU (``torchrun``).  Frames are independent given the camera
parameters, so each rank owns a contiguous range of frames, runs the fused
residual/Jacobian/Schur kernels on its range, and only the packed reduced camera
system ``[S | b | g | diag | scalars]`` (5.4k doubles at 6 cameras) is summed
once per evaluation inside ``libmcba`` (one kernel over NVLink peer memory, or
NCCL).  Every rank solves the small system redundantly (identical inputs, so no
broadcast) and back-substitutes its own poses.  The front end of ``bundle_adjust``
(bundle_adjustment.py:265-296) is sharded the same way: every rank uploads and
scans only ITS frames, the global nanmedian of the per-point errors is found by an
exact radix selection whose 256-bin histograms are the only thing summed across
ranks, and rank 0 draws the random sub-sample for everybody.  ``torch.distributed``
provides the rendezvous and those small sums; the data path has no collective
other than the reduced-system sum.
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
    """All ranks receive the list of every rank's 1-D array (float64 / int64; poses, gradients, frame
    indices).  NCCL groups: one padded ``all_gather_into_tensor`` over NVLink (no pickling); gloo
    groups (CPU tests): host objects."""
    a = np.asarray(local)
    if world_size() == 1:
        return [a]
    dist = _dist()
    W = world_size()
    if dist.get_backend() != "nccl" or a.dtype not in (np.float64, np.int64):
        out = [None] * W
        dist.all_gather_object(out, a)
        return out
    import torch
    flat = np.ascontiguousarray(a).ravel()
    n = allreduce_sum(np.eye(W, dtype=np.int64)[rank()] * flat.size)
    cap = max(int(n.max()), 1)
    mine = torch.zeros(cap, dtype=torch.float64 if a.dtype == np.float64 else torch.int64, device="cuda")
    mine[:flat.size] = torch.from_numpy(flat).cuda()
    full = torch.empty(W * cap, dtype=mine.dtype, device="cuda")
    dist.all_gather_into_tensor(full, mine)
    host = full.cpu().numpy()
    return [host[r * cap: r * cap + int(n[r])].copy() for r in range(W)]


def allreduce_sum(values):
    """Element-wise sum over ranks of a small host array (int64 / float64); every rank gets the
    result.  NCCL groups carry it in a device tensor, gloo groups (CPU tests) in a host tensor."""
    import torch
    dist = _dist()
    a = np.ascontiguousarray(values)
    if world_size() == 1:
        return a.copy()
    t = torch.from_numpy(a.copy())
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t)
    return t.cpu().numpy()


def broadcast_object(obj, src=0):
    dist = _dist()
    if world_size() == 1:
        return obj
    box = [obj if rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def kth_smallest(local_histogram, ks, reduce=allreduce_sum):
    """Exact ``k``-th smallest (0-based, one per entry of ``ks``) of the union over ranks of
    non-negative float64 values, by most-significant-digit radix selection on the IEEE-754 bit
    patterns: eight passes of 8 bits; per pass every rank histograms ITS values under the
    current prefixes (``local_histogram(prefix, prefix_bits) -> 256 int64 counts``) and the counts
    are summed over ranks.  Nothing but 256-entry histograms crosses ranks."""
    state = [(0, int(k)) for k in ks]          # (prefix, rank of the wanted value among the values under the prefix)
    for p in range(8):
        prefixes = sorted({pre for pre, _ in state})
        local = np.stack([np.asarray(local_histogram(pre, 8 * p), dtype=np.int64) for pre in prefixes])
        total = reduce(local)
        nxt = []
        for pre, k in state:
            cum = np.cumsum(total[prefixes.index(pre)])
            b = int(np.searchsorted(cum, k, side="right"))
            if b > 255:
                raise ValueError("kth_smallest: k is not smaller than the number of values")
            nxt.append(((pre << 8) | b, k - (int(cum[b - 1]) if b else 0)))
        state = nxt
    return [float(np.array([pre], dtype=np.uint64).view(np.float64)[0]) for pre, _ in state]


def global_nanmedian(local_histogram, n_finite_total, reduce=allreduce_sum):
    """``np.nanmedian`` of the union over ranks (mean of the two middle values for an even count)."""
    n = int(n_finite_total)
    if n == 0:
        return float("nan")
    if n & 1:
        return kth_smallest(local_histogram, [n // 2], reduce)[0]
    lo, hi = kth_smallest(local_histogram, [n // 2 - 1, n // 2], reduce)
    return 0.5 * (lo + hi)


def gather_device_vectors(local):
    """Concatenation inputs for a per-rank device vector: returns the list of every rank's vector
    (device tensors on NCCL groups: one padded ``all_gather_into_tensor`` over NVLink)."""
    import torch
    dist = _dist()
    W = world_size()
    n = allreduce_sum(np.eye(W, dtype=np.int64)[rank()] * int(local.numel()))
    if dist.get_backend() != "nccl":
        out = [None] * W
        dist.all_gather_object(out, local.cpu().numpy())
        return [torch.from_numpy(o) for o in out]
    cap = int(n.max())
    mine = torch.zeros(cap, dtype=local.dtype, device=local.device)
    mine[:local.numel()] = local
    full = torch.empty(W * cap, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(full, mine)
    return [full[r * cap: r * cap + int(n[r])] for r in range(W)]


def gather_concat(vectors):
    """ONE exchange for several per-rank vectors.  ``vectors``: the same number of 1-D float64 vectors on
    every rank (CUDA tensors or numpy arrays, any lengths).  Returns, per vector, the list of every
    rank's piece (CUDA tensors on NCCL groups -- one sum of the length table and one padded
    ``all_gather_into_tensor`` over NVLink -- host tensors on gloo groups)."""
    import torch
    dist = _dist()
    W, r = world_size(), rank()
    if W == 1:
        return [[v if hasattr(v, "cpu") else torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64))] for v in vectors]
    if dist.get_backend() != "nccl":
        host = [v.cpu().numpy() if hasattr(v, "cpu") else np.asarray(v, dtype=np.float64) for v in vectors]
        out = [None] * W
        dist.all_gather_object(out, host)
        return [[torch.from_numpy(np.ascontiguousarray(out[q][k])) for q in range(W)] for k in range(len(vectors))]
    dev = torch.device("cuda", torch.cuda.current_device())
    vs = [v if hasattr(v, "is_cuda") else torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(dev) for v in vectors]
    lens = np.zeros((W, len(vs)), dtype=np.int64)
    lens[r] = [int(v.numel()) for v in vs]
    lens = allreduce_sum(lens)
    cap = max(int(lens.sum(axis=1).max()), 1)
    mine = torch.zeros(cap, dtype=torch.float64, device=dev)
    torch.cat([v.reshape(-1).to(torch.float64) for v in vs], out=mine[:int(lens[r].sum())])
    full = torch.empty(W * cap, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(full, mine)
    offs = np.concatenate([np.zeros((W, 1), dtype=np.int64), np.cumsum(lens, axis=1)], axis=1)
    return [[full[q * cap + int(offs[q, k]): q * cap + int(offs[q, k + 1])] for q in range(W)] for k in range(len(vs))]


_sharded = {}


def _sharded_problem(uvs_local, calib_objpoints, dev, counts):
    """The rank's device problem WITH its communicator (NCCL rendezvous + CUDA IPC peer buffers,
    hundreds of milliseconds to set up) is kept between calls whose per-rank frame counts are the
    same on every rank; the key is the full count vector, so all ranks hit or miss together and the
    collective set-up below stays collective."""
    from .engine import BAProblem
    key = (int(uvs_local.shape[0]), int(uvs_local.shape[2]), tuple(int(c) for c in counts), int(dev))
    prob = _sharded.get(key)
    if prob is None:
        release_sharded_problem()
        uid = broadcast_unique_id()
        prob = BAProblem(uvs_local, calib_objpoints, device=dev, comm=(uid, rank(), world_size()))
        _sharded[key] = prob
    else:
        prob.set_observations(uvs_local, calib_objpoints)
    return prob


def release_sharded_problem():
    """Free the cached sharded problem and its communicator (call on every rank)."""
    for old in _sharded.values():
        old.close()
    _sharded.clear()


def solve_sharded(uvs_local, calib_objpoints, x0_local, frames_per_rank=None, **opt_kwargs):
    """Solve with frames sharded over the ranks of the default process group.

    ``uvs_local`` (C,F_r,N,2): THIS rank's frames (numpy array or float64 CUDA tensor);
    ``x0_local``: the 12C camera parameters (identical on all ranks) followed by the poses of this
    rank's frames.  Ranks may hold different numbers of frames (at least one each;
    ``frames_per_rank``: the counts of all ranks when the caller already knows them).  Returns
    ``(x, result)``: the full solution, poses concatenated in rank order, on every rank;
    ``result.fun`` is the full residual vector in the reference's order
    (bundle_adjustment.py:97: camera-major, so every camera's block is the concatenation of the
    ranks' blocks), gathered over NVLink at the end of the solve -- together with the poses and the
    gradient, in one exchange -- and copied to the host on first access.
    """
    import torch
    W, r = world_size(), rank()
    C, F_local = int(uvs_local.shape[0]), int(uvs_local.shape[1])
    counts = (np.asarray(frames_per_rank, dtype=np.int64) if frames_per_rank is not None
              else allreduce_sum(np.eye(W, dtype=np.int64)[r] * F_local))
    if counts.shape != (W,) or int(counts[r]) != F_local:
        raise ValueError("frames_per_rank does not match this rank's frames")
    if counts.min() < 1:
        raise ValueError(f"every rank needs at least one frame (frames per rank: {counts.tolist()})")
    start = int(counts[:r].sum())
    dev = local_device()
    torch.cuda.set_device(dev)
    prob = _sharded_problem(uvs_local, calib_objpoints, dev, counts)
    nc = 12 * C
    try:
        x_loc, result = prob.solve(np.asarray(x0_local, dtype=np.float64), **opt_kwargs)
        result["peer_memory"] = bool(getattr(prob, "peer_memory", False))
        result["collective"] = "peer" if result["peer_memory"] else "nccl"
        # residuals at the solution (computed by every rank on its frames), poses and pose gradients:
        # one exchange; per_cam = residual scalars per camera, where a rank's block goes in the full vector
        d_r, per_cam = prob.residuals_device(x_loc)
        parts, poses, grads, per_cam_all = gather_concat([d_r, x_loc[nc:], result.grad[nc:], per_cam.astype(np.float64)])
        per_cam_all = np.stack([p.cpu().numpy() for p in per_cam_all]).astype(np.int64)      # (W, C)
    except BaseException:
        release_sharded_problem()     # a failed collective leaves the communicator unusable
        raise
    x = merge_params(x_loc[:nc], [p.cpu().numpy() for p in poses])
    result["x"] = x
    result["grad"] = merge_params(result.grad[:nc], [g.cpu().numpy() for g in grads])
    result["active_mask"] = np.zeros_like(x)
    result["shard"] = (start, start + F_local)

    def fun():
        offs = np.concatenate([np.zeros((W, 1), dtype=np.int64), np.cumsum(per_cam_all, axis=1)], axis=1)
        pieces = [parts[q][int(offs[q, c]): int(offs[q, c + 1])] for c in range(C) for q in range(W)]
        full = torch.cat(pieces) if pieces else parts[0][:0]
        from . import _native
        return _native.to_host(full) if full.is_cuda else full.numpy()
    result.set_lazy("fun", fun)
    return x, result
