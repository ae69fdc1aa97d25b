"""SURVEY 8(f) rows N1 / N2 at BASELINE configs[2] shapes (6 cameras x 50,000 frames): the device
front end of bundle_adjust (select_frames) and the initialisation algebra (estimate_all_extrinsics,
consensus_calib_poses) through the public API with numpy arrays in and out (wall clock, uploads
included), beside the oracle port (the reference's numpy code path) on the same inputs.
Prints one JSON object; `python scripts/bench_init.py > profiles/rNN_init.json`."""
import contextlib, io, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import multicam_calibration_b200 as mcc
from multicam_calibration_b200.synthetic import make_scene, make_camera_poses
from oracle import np_oracle as orc

C = 6
F = int(sys.argv[1]) if len(sys.argv) > 1 else 50000


def wall(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best, r


def cpu(fn):
    t0 = time.perf_counter()
    r = fn()
    return time.perf_counter() - t0, r


out = {"cameras": C, "frames": F, "host_cores": os.cpu_count()}
poses, ext_true, board_true = make_camera_poses(C, F, p_missing=0.2, seed=0)
t_gpu, (ext, tree) = wall(lambda: mcc.estimate_all_extrinsics(poses))
t_cpu, (ext_o, tree_o) = cpu(lambda: orc.estimate_all_extrinsics(poses))
out["estimate_all_extrinsics"] = {"api_ms": t_gpu * 1e3, "cpu_port_ms": t_cpu * 1e3, "speedup": t_cpu / t_gpu,
                                  "max_abs_diff_vs_oracle": float(np.nanmax(np.abs(ext - ext_o))),
                                  "h2d_bytes": poses.nbytes}
t_gpu, cons = wall(lambda: mcc.consensus_calib_poses(poses, ext))
t_cpu, cons_o = cpu(lambda: orc.consensus_calib_poses(poses, ext))
out["consensus_calib_poses"] = {"api_ms": t_gpu * 1e3, "cpu_port_ms": t_cpu * 1e3, "speedup": t_cpu / t_gpu,
                                "max_abs_diff_vs_oracle": float(np.nanmax(np.abs(cons - cons_o))),
                                "h2d_bytes": poses.nbytes, "d2h_bytes": cons.nbytes}

sc = make_scene(C, F, sigma=0.5, p_missing_view=0.2, seed=0)
args = sc.init_args()
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    t_gpu, use = wall(lambda: mcc.select_frames(*args, n_frames=None))
    t_cpu, (use_o, thr_o) = cpu(lambda: orc.select_frames(*args, None, None))
out["select_frames"] = {"api_ms": t_gpu * 1e3, "cpu_port_ms": t_cpu * 1e3, "speedup": t_cpu / t_gpu,
                        "identical_frames": bool(np.array_equal(use, use_o)), "h2d_bytes": sc.uvs.nbytes}
print(json.dumps(out, indent=1))
