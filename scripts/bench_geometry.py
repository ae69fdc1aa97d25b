"""BASELINE configs[4] (batched geometry.triangulate + project_points of 1M keypoints across 6
cameras) and the K1 residual / predict kernels at configs[2], measured on the device with CUDA
events through the C ABI (inputs resident in HBM), beside the oracle port on a bounded sample.
Prints one JSON object; `python scripts/bench_geometry.py > profiles/rNN_geometry.json`."""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import multicam_calibration_b200 as mcc
from multicam_calibration_b200 import _native
from multicam_calibration_b200.synthetic import make_scene
from oracle import np_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
lib = _native.load()
dev = 0
torch.cuda.set_device(dev)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
out = {"points": P, "hbm_peak_gbs": peaks["hbm_gbs"]}


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def vp(a):
    return np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(ctypes.c_void_p)


# ---------------------------------------------------------------- config 5
sc = make_scene(6, 8, sigma=0.0, seed=0)
cams = sc.true_cams
rng = np.random.default_rng(0)
pts = rng.normal(0, 80, (P, 3))
ext = cams[:, 6:12]
Ks = np.zeros((6, 3, 3)); Ks[:, 0, 0] = cams[:, 0]; Ks[:, 1, 1] = cams[:, 1]; Ks[:, 0, 2] = cams[:, 2]; Ks[:, 1, 2] = cams[:, 3]; Ks[:, 2, 2] = 1
dist = np.zeros((6, 5)); dist[:, :2] = cams[:, 4:6]
stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
d_pts = torch.as_tensor(pts).cuda()
d_uv = torch.empty((6, P, 2), dtype=torch.float64, device="cuda")


def project_all():
    for c in range(6):
        _native.check(lib.mcba_project_points(dev, stream, ctypes.c_void_p(d_pts.data_ptr()), P, vp(ext[c]), vp(Ks[c]),
                                              vp(dist[c, :2]), ctypes.c_void_p(d_uv[c].data_ptr())))


def project_multi():
    _native.check(lib.mcba_project_points_multi(dev, stream, ctypes.c_void_p(d_pts.data_ptr()), P, 6, vp(ext), vp(Ks),
                                                vp(dist[:, :2]), ctypes.c_void_p(d_uv.data_ptr())))


ms_m = timed(project_multi)
out["project_points_multi"] = {"ms": ms_m, "views_per_s": 6 * P / (ms_m * 1e-3), "algorithmic_bytes": P * 24 + 6 * P * 16,
                               "achieved_gbs": (P * 24 + 6 * P * 16) / (ms_m * 1e-3) / 1e9,
                               "hbm_frac": (P * 24 + 6 * P * 16) / (ms_m * 1e-3) / 1e9 / peaks["hbm_gbs"],
                               "note": "one pass over the points for all 6 cameras"}
ms = timed(project_all)
out["project_points"] = {"ms": ms, "views_per_s": 6 * P / (ms * 1e-3), "algorithmic_bytes": 6 * P * (24 + 16),
                         "achieved_gbs": 6 * P * 40 / (ms * 1e-3) / 1e9}
out["project_points"]["hbm_frac"] = out["project_points"]["achieved_gbs"] / peaks["hbm_gbs"]
uv = d_uv.cpu().numpy() + rng.normal(0, 0.3, (6, P, 2))
uv[rng.random((6, P)) < 0.2] = np.nan
d_obs = torch.as_tensor(uv).cuda()
d_out = torch.empty((P, 3), dtype=torch.float64, device="cuda")


def tri():
    _native.check(lib.mcba_triangulate(dev, stream, ctypes.c_void_p(d_obs.data_ptr()), 6, P, vp(ext), vp(Ks), vp(dist),
                                       ctypes.c_void_p(d_out.data_ptr())))


ms = timed(tri)
out["triangulate"] = {"ms": ms, "points_per_s": P / (ms * 1e-3), "algorithmic_bytes": P * (6 * 16 + 24),
                      "achieved_gbs": P * 120 / (ms * 1e-3) / 1e9}
out["triangulate"]["hbm_frac"] = out["triangulate"]["achieved_gbs"] / peaks["hbm_gbs"]
res = d_out.cpu().numpy()
ok = np.isfinite(res).all(1)
out["triangulate"]["median_error_vs_truth"] = float(np.median(np.linalg.norm(res[ok] - pts[ok], axis=1)))
out["triangulate"]["fraction_triangulated"] = float(ok.mean())
# CPU baseline: oracle port (numpy restatement of geometry.triangulate incl. the per-point nanmedian loop) on 5k points
ns = 5000
intr = [(Ks[c], dist[c]) for c in range(6)]
t0 = time.perf_counter()
ref = orc.triangulate([uv[c, :ns] for c in range(6)], list(ext), intr)
dt = time.perf_counter() - t0
out["triangulate"]["cpu_baseline"] = {"points_per_s": ns / dt, "kind": "port", "cores": 1, "sample": f"{ns} points x 6 views, {dt:.2f} s"}
good = np.isfinite(ref).all(1) & ok[:ns]
out["triangulate"]["max_abs_diff_vs_oracle_sample"] = float(np.abs(res[:ns][good] - ref[good]).max())

# ---------------------------------------------------------------- K1 at configs[2]
sc3 = make_scene(6, 50000, sigma=0.5, p_missing_view=0.2, seed=0)
prob = mcc.BAProblem(sc3.uvs, sc3.objpoints, device=dev)
h = prob._h
with torch.cuda.stream(prob.stream):
    d_x = torch.as_tensor(sc3.x0()).cuda()
    m = prob.n_residuals
    d_r = torch.empty(m, dtype=torch.float64, device="cuda")
    d_p = torch.empty(sc3.uvs.shape, dtype=torch.float64, device="cuda")
    ms_r = timed(lambda: _native.check(lib.mcba_residuals(h, ctypes.c_void_p(d_x.data_ptr()), ctypes.c_void_p(d_r.data_ptr()))))
    ms_p = timed(lambda: _native.check(lib.mcba_predict(h, ctypes.c_void_p(d_x.data_ptr()), ctypes.c_void_p(d_p.data_ptr()))))
n_obs = sc3.n_obs
slots = int(np.prod(sc3.uvs.shape[:3]))
out["k1_residuals"] = {"ms": ms_r, "obs_per_s": n_obs / (ms_r * 1e-3), "algorithmic_bytes": 16 * slots + 8 * m,
                       "achieved_gbs": (16 * slots + 8 * m) / (ms_r * 1e-3) / 1e9,
                       "note": "reads every (c,f,n) slot incl. NaN ones (16 B), writes m compacted scalars (8 B)"}
out["k1_residuals"]["hbm_frac"] = out["k1_residuals"]["achieved_gbs"] / peaks["hbm_gbs"]
out["k1_predict"] = {"ms": ms_p, "slots_per_s": slots / (ms_p * 1e-3), "algorithmic_bytes": 16 * slots,
                     "achieved_gbs": 16 * slots / (ms_p * 1e-3) / 1e9}
out["k1_predict"]["hbm_frac"] = out["k1_predict"]["achieved_gbs"] / peaks["hbm_gbs"]
print(json.dumps(out, indent=1))
