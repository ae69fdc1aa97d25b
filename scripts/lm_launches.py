"""One BA solve at BASELINE configs[2] (after a warm-up solve) for an ncu launch list:

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lm.csv \
        python scripts/lm_launches.py [cams frames]

The last `mcba_lm_run` in the list is the measured solve (reference defaults)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multicam_calibration_b200 as mcc
from multicam_calibration_b200.synthetic import make_scene

C = int(sys.argv[1]) if len(sys.argv) > 1 else 6
F = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
sc = make_scene(C, F, sigma=0.5, p_missing_view=0.2, seed=0)
prob = mcc.BAProblem(sc.uvs, sc.objpoints)
x0 = sc.x0()
prob.solve(x0, verbose=0, max_nfev=3)
x, res = prob.solve(x0, verbose=0)
print(f"iterations {res.iterations} nfev {res.nfev} device_ms {res.solve_ms:.3f} launches {res.kernel_launches}")
