"""Host-buffer call (mcba_build_reduced_host) at BASELINE configs[2]: pipelined against plain, page-locked buffers."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import multicam_calibration_b200 as mcc
from multicam_calibration_b200 import _native
from multicam_calibration_b200.synthetic import make_scene
F = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
sc = make_scene(6, F, sigma=0.5, p_missing_view=0.2, seed=0)
x0 = sc.x0(); C = 6
lib = _native.load()
prob = mcc.BAProblem(sc.uvs, sc.objpoints)
keep = [torch.from_numpy(np.ascontiguousarray(sc.uvs)).pin_memory(), torch.from_numpy(np.ascontiguousarray(sc.objpoints, dtype=np.float64)).pin_memory(),
        torch.from_numpy(x0.copy()).pin_memory(), torch.empty(144 * C * C, dtype=torch.float64).pin_memory(),
        torch.empty(12 * C, dtype=torch.float64).pin_memory(), torch.empty(1, dtype=torch.float64).pin_memory()]
ptrs = [ctypes.c_void_p(t.data_ptr()) for t in keep]
def call():
    _native.check(lib.mcba_build_reduced_host(prob._h, ptrs[0], ptrs[1], ptrs[2], 1e-3, 1, 1.0, ptrs[3], ptrs[4], ptrs[5]))
for mode in ("pipelined", "plain", "pipelined", "plain"):
    if mode == "plain": os.environ["MCBA_NO_HOST_PIPELINE"] = "1"
    else: os.environ.pop("MCBA_NO_HOST_PIPELINE", None)
    for _ in range(3): call()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): call()
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 100
    print(f"{mode}: {ms:.3f} ms per call  cost {float(keep[5][0]):.6f}")
# raw copy rates for reference
d = torch.empty_like(keep[0], device="cuda")
for _ in range(2): d.copy_(keep[0], non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): d.copy_(keep[0], non_blocking=True)
torch.cuda.synchronize(); print(f"one contiguous H2D of {keep[0].numel()*8/1e6:.0f} MB: {(time.perf_counter()-t0)*200:.3f} ms")
