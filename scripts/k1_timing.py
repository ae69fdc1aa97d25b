"""K1 (residual_chunks_kernel) timing at BASELINE configs[2]: `python scripts/k1_timing.py` prints ms per call and
the fraction of the measured HBM peak by SURVEY 8(d) bytes (32 B per finite observation + 48 B per frame)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multicam_calibration_b200 as mcc
from multicam_calibration_b200 import _native
from multicam_calibration_b200.synthetic import make_scene

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) else 6650.0
sc = make_scene(6, 50000, sigma=0.5, p_missing_view=0.2, seed=0)
prob = mcc.BAProblem(sc.uvs, sc.objpoints)
lib = _native.load()
with torch.cuda.stream(prob.stream):
    d_x = torch.as_tensor(sc.x0()).cuda()
    d_r = torch.empty(prob.n_residuals, dtype=torch.float64, device="cuda")
    call = lambda: _native.check(lib.mcba_residuals(prob._h, ctypes.c_void_p(d_x.data_ptr()), ctypes.c_void_p(d_r.data_ptr())))
    for _ in range(5):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(prob.stream)
    for _ in range(50):
        call()
    e1.record(prob.stream)
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
alg = 32.0 * sc.n_obs + 48.0 * sc.uvs.shape[1]
print(f"K1 {ms:.4f} ms  {alg / ms / 1e6:.0f} GB/s  frac {alg / ms / 1e6 / peak:.3f}")
