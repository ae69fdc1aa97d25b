"""Wall time of the public bundle_adjust(...) call at BASELINE configs[2] (6 cams x 50k frames,
20 % missing views): device front end + upload + solve, as a user of the reference API sees it."""
import io, contextlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import multicam_calibration_b200 as mcc
from multicam_calibration_b200.synthetic import make_scene

F = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
sc = make_scene(6, F, sigma=0.5, p_missing_view=0.2, seed=0)
args = sc.init_args()
for rep in range(3):
    np.random.seed(0)
    buf = io.StringIO()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(buf):
        t1 = time.perf_counter()
        use = mcc.select_frames(*args, n_frames=None)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        np.random.seed(0)
        e, i, p, use, res = mcc.bundle_adjust(*args, n_frames=None, verbose=0)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"rep {rep}: select_frames {1e3*(t2-t1):.1f} ms | bundle_adjust total {1e3*(t3-t2):.1f} ms "
          f"(device solve {res.solve_ms:.2f} ms, {res.iterations} it, rms {res.rms:.4f} px, {len(use)} frames) "
          f"| {buf.getvalue().splitlines()[0]}")
