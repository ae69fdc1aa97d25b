"""CPU: host-side logic, the C-ABI surface and the multi-rank plumbing (gloo)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden, split_cams
from oracle import np_oracle as orc
import multicam_calibration_b200 as mcc
from multicam_calibration_b200 import _native, distributed, engine


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "mcba.h")).read()
    declared = set(re.findall(r"\b(mcba_[a-z_0-9]+)\s*\(", header))
    declared -= {"mcba_handle", "mcba_options", "mcba_result"}
    assert declared, "no prototypes found"
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/mcba.h but not exported"
    assert declared == set(_native.EXPORTED_SYMBOLS), declared ^ set(_native.EXPORTED_SYMBOLS)
    assert _native.load().mcba_version() >= 100


def test_default_options_follow_the_reference_defaults():
    o = _native.Options()
    _native.load().mcba_default_options(ctypes.byref(o))
    assert (o.ftol, o.xtol, o.gtol, o.loss, o.verbose) == (1e-4, 1e-8, 1e-8, 1, 2)   # bundle_adjustment.py:301-303
    o2 = engine.parse_options(dict(ftol=1e-9, loss="linear", max_nfev=7, verbose=0), 100)
    assert (o2.ftol, o2.loss, o2.max_nfev, o2.verbose) == (1e-9, 0, 7, 0)
    for bad in (dict(jac="3-point"), dict(tr_solver="exact"), dict(bounds=(0, 1))):
        with pytest.raises(TypeError):
            engine.parse_options(bad, 10)
    with pytest.raises(ValueError):
        engine.parse_options(dict(loss="huber"), 10)
    with pytest.raises(ValueError):
        engine.parse_options(dict(x_scale=1.0), 10)


def test_no_cuda_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    g = load_golden("ba_small")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mcc.residuals(g["x0"], g["uvs"], g["objpoints"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mcc.project_points(np.zeros((3, 3)), np.zeros(6), np.eye(3))


def test_serialize_roundtrip_and_layout():
    g = load_golden("ba_small")
    x0 = g["x0"]
    C = g["uvs"].shape[0]
    ext, intr, poses = mcc.deserialize_params(x0, C)
    ext_o, intr_o, poses_o = orc.deserialize_params(x0, C)
    assert np.array_equal(ext, ext_o) and np.array_equal(poses, poses_o)
    for (K, d), (Ko, do) in zip(intr, intr_o):
        assert np.array_equal(K, Ko) and np.array_equal(d, do) and d.shape == (5,)
    assert np.array_equal(mcc.serialize_params(ext, intr, poses), x0)


def test_sparsity_pattern_matches_reference():
    g = load_golden("ba_small")
    A = mcc.bundle_adjustment_sparsity(g["uvs"]).tocsr()
    A.sort_indices()
    assert np.array_equal(A.indices, g["A_indices"]) and np.array_equal(A.indptr, g["A_indptr"])
    assert (A.sum(1) == 18).all()


def test_host_geometry_helpers_match_reference():
    g = load_golden("geometry")
    np.testing.assert_allclose(mcc.rodrigues(g["r"]), g["R"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(mcc.rodrigues_inv(g["R"]), g["r_back"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(mcc.get_transformation_matrix(g["t6"]), g["T"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(mcc.get_transformation_vector(g["T"]), g["t6_back"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(mcc.apply_rigid_transform(g["ext"], g["pts"]), g["pts_rigid_vec"], rtol=1e-14)
    np.testing.assert_allclose(mcc.apply_rigid_transform(g["T"][3], g["pts"][3]),
                               orc.apply_rigid_transform(g["T"][3], g["pts"][3]), rtol=1e-14)
    np.testing.assert_allclose(mcc.get_projection_matrix(g["ext"], (g["K"], g["dist"])), g["P"], rtol=1e-14)
    h = mcc.euclidean_to_homogenous(g["pts"])
    assert h.shape[-1] == 4 and (h[..., 3] == 1).all()
    np.testing.assert_allclose(mcc.homogeneous_to_euclidean(h * 3.0), g["pts"], rtol=1e-15)
    t, rmsd = mcc.rigid_transform_from_correspondences(g["pts"][0], g["pts_rigid_vec"][0])
    np.testing.assert_allclose(t, g["ext"], atol=1e-9)
    assert rmsd < 1e-9


def test_shard_bounds_partition_frames():
    for F, W in ((10, 3), (50000, 8), (8, 8), (33, 2)):
        spans = [distributed.shard_bounds(F, W, r) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == F
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    x = np.arange(12 * 2 + 6 * 7, dtype=float)
    parts = [distributed.split_params(x, 2, *distributed.shard_bounds(7, 3, r)) for r in range(3)]
    assert np.array_equal(distributed.merge_params(parts[0][:24], [p[24:] for p in parts]), x)


def _np_histogrammer(vals):
    """numpy restatement of mcba_key_histogram (csrc/k0_frontend.cu) for one rank's values."""
    keys = np.ascontiguousarray(vals[~np.isnan(vals)]).view(np.uint64)

    def hist(prefix, bits):
        k = keys if bits == 0 else keys[(keys >> np.uint64(64 - bits)) == np.uint64(prefix)]
        return np.bincount(((k >> np.uint64(64 - bits - 8)) & np.uint64(0xff)).astype(np.int64), minlength=256)
    return hist


def test_radix_selection_gives_the_exact_global_nanmedian():
    """distributed.global_nanmedian (the sharded front end's replacement for np.nanmedian over all
    ranks' per-point errors, bundle_adjustment.py:281-282) is exact: odd / even counts, ties, zeros,
    subnormals, NaNs, one rank without any finite value."""
    rng = np.random.default_rng(0)
    for n, shards in ((1, 1), (2, 2), (1001, 3), (4096, 4), (20000, 8)):
        vals = np.abs(rng.normal(0, 3.0, n)) ** rng.integers(1, 4, n)
        vals[rng.random(n) < 0.2] = np.nan
        vals[rng.integers(0, n, n // 10)] = vals[0] if vals[0] == vals[0] else 1.5      # ties
        vals[rng.integers(0, n, 3)] = 0.0
        vals[rng.integers(0, n, 2)] = 5e-324
        if np.isnan(vals).all():
            vals[0] = 2.0
        parts = np.array_split(vals, shards)
        if shards > 2:
            parts[1] = np.full_like(parts[1], np.nan)
        hists = [_np_histogrammer(p) for p in parts]
        merged = np.concatenate(parts)
        local = lambda prefix, bits: np.sum([h(prefix, bits) for h in hists], axis=0)
        n_finite = int((~np.isnan(merged)).sum())
        got = distributed.global_nanmedian(local, n_finite, reduce=lambda a: a)
        assert got == np.nanmedian(merged), (n, shards)
        ks = sorted({0, n_finite // 3, n_finite - 1})
        assert distributed.kth_smallest(local, ks, reduce=lambda a: a) == list(np.sort(merged[~np.isnan(merged)])[ks])
    assert np.isnan(distributed.global_nanmedian(lambda p, b: np.zeros(256, dtype=np.int64), 0, reduce=lambda a: a))
    with pytest.raises(ValueError):
        distributed.kth_smallest(_np_histogrammer(np.array([1.0, 2.0])), [2], reduce=lambda a: a)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch
        g = load_golden("ba_small")
        uvs, obj, x0 = g["uvs"], g["objpoints"], g["x0"]
        C, F = uvs.shape[:2]
        a, b = distributed.shard_bounds(F, world, rank)
        assert distributed.world_size() == world and distributed.rank() == rank
        # each rank's reduced camera system (oracle arithmetic) summed over ranks == the full one:
        # the camera block U is additive over frames, and so are the Schur corrections
        xl = distributed.split_params(x0, C, a, b)
        H, grad, cost = orc.normal_equations(xl, uvs[:, a:b], obj)
        S, bb = orc.reduced_camera_system(H, grad, C)
        packed = torch.from_numpy(np.concatenate([S.ravel(), bb, [cost]]))
        dist.all_reduce(packed)
        poses = distributed.gather_arrays(xl[12 * C:])
        x_back = distributed.merge_params(xl[:12 * C], poses)
        # the small host-side collectives of the sharded front end: counters, the broadcast of rank 0's
        # random sub-sample (ranks have DIFFERENT RNG states) and the exact global median
        assert np.array_equal(distributed.allreduce_sum(np.array([rank + 1, 10], dtype=np.int64)), [3, 20])
        np.random.seed(100 + rank)
        chosen = distributed.broadcast_object(np.random.choice(50, 7, replace=False) if rank == 0 else None)
        np.random.seed(100)
        assert np.array_equal(chosen, np.random.choice(50, 7, replace=False))
        err = np.abs(np.random.default_rng(3).normal(0, 2, 999))
        err[::7] = np.nan
        mine = err[rank::world]
        med = distributed.global_nanmedian(_np_histogrammer(mine), distributed.allreduce_sum(
            np.array([(~np.isnan(mine)).sum()], dtype=np.int64))[0])
        assert med == np.nanmedian(err)
        vec = distributed.gather_device_vectors(torch.arange(3 + rank, dtype=torch.float64) + 10 * rank)
        assert [v.tolist() for v in vec] == [[0.0, 1.0, 2.0], [10.0, 11.0, 12.0, 13.0]]
        # several per-rank vectors in one exchange (residuals + poses + gradients of the sharded solve)
        a_, b_ = distributed.gather_concat([torch.arange(2 + rank, dtype=torch.float64), np.array([float(rank)])])
        assert [v.tolist() for v in a_] == [[0.0, 1.0], [0.0, 1.0, 2.0]] and [v.tolist() for v in b_] == [[0.0], [1.0]]
        if rank == 0:
            Hf, gf, cf = orc.normal_equations(x0, uvs, obj)
            Sf, bf = orc.reduced_camera_system(Hf, gf, C)
            full = np.concatenate([Sf.ravel(), bf, [cf]])
            err = np.abs(packed.numpy() - full).max() / np.abs(full).max()
            q.put((err, bool(np.array_equal(x_back, x0))))
    finally:
        dist.destroy_process_group()


def test_frame_sharding_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, same = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert err < 1e-12 and same


def test_optimize_result_lazy_fields():
    """result.fun-style lazy fields: computed once on first access, visible in keys(), never stored as dict items."""
    calls = []
    r = engine.OptimizeResult(x=1, cost=2.0)
    r.set_lazy("fun", lambda: calls.append(1) or [1, 2, 3])
    assert "fun" in r and "fun" in r.keys() and not dict.__contains__(r, "fun")
    assert set(dict(r)) == {"x", "cost"}
    assert r.fun == [1, 2, 3] and r["fun"] == [1, 2, 3] and calls == [1]
    assert dict.__contains__(r, "fun")
    assert r.get("missing", 5) == 5 and not hasattr(r, "missing")
    with pytest.raises(AttributeError):      # scipy's OptimizeResult raises too: a typo is not None
        r.missing
    with pytest.raises(KeyError):
        r["missing"]
    t = engine.OptimizeResult(a=1)
    t.set_lazy("fun", lambda: 7)
    assert t.pop_lazy("fun")() == 7 and "fun" not in t


def _run_bench(*argv, env=None):
    import subprocess
    import sys
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], capture_output=True, text=True,
                          timeout=600, env=e)


def test_bench_reference_arm_prints_one_contract_line():
    """bench.py --impl reference runs on host cores only and prints exactly one JSON line with the
    contract's keys (small samples here; the defaults are sized for the GPU box's host)."""
    import json
    out = _run_bench("--impl", "reference", "--steps", "1", "--warmup", "0",
                     env={"MCBA_BENCH_SAMPLE_FRAMES": "60", "MCBA_BENCH_CONVERGE_FRAMES": "24"})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "obs/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("obs/s residual+Jacobian+Schur") and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["vs_baseline"] is None
    assert "workload" in d["config"] and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and cb["sample"]
    assert cb["converge"]["status"] > 0 and cb["converge"]["wall_s"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "obs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_engine_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = _run_bench("--steps", "1", "--warmup", "0")
    assert out.returncode != 0
    assert "no CPU fallback" in out.stderr and not out.stdout.strip()
